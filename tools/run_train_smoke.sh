#!/bin/bash
# tiny synthetic run of the reference-shaped entry point (2 epochs x 6 iterations)
cd "$(dirname "$0")/../sc-sfmlearner-release_b200"
python train.py synthetic --name smoke --with-pretrain 0 --epochs 2 --epoch-size 6 -b 2 --print-freq 2 --with-auto-mask 1 --synthetic-size 128 416 "$@"
