"""Stand-in for the two matplotlib names the reference's utils.py imports (colour maps for tensorboard images only)."""
