"""Launched by tests/test_ddp_gpu.py under torchrun (2 ranks, one GPU each): the data-parallel step against the oracle.

  mode "ddp"    default semantics: every rank is the reference at batch B/N; the all-reduced gradient arena must equal the MEAN
                of the per-shard oracle gradients (SURVEY.md section 4 iii / 8e).
  mode "exact"  Trainer(exact_global_masks=True): per-rank BatchNorm + losses normalised over the GLOBAL batch = what the
                reference's nn.DataParallel computes (train.py:168-169; forward per GPU slice, losses on the gathered outputs).
                Oracle: both shards through the same oracle networks, outputs concatenated, losses on the full batch.
  also: the captured (CUDA-graph) data-parallel step reproduces the eager one, replicas stay bit-identical.
Prints "DDP_CHECK_OK <mode>" per passed mode on rank 0; any failure raises.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "sc-sfmlearner-release_b200"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import models  # noqa: E402
from golden_util import det_weights  # noqa: E402
from oracle import losses as OL  # noqa: E402
from oracle import nets as N  # noqa: E402
from oracle import step as OS  # noqa: E402
from scsfm import synth  # noqa: E402
from scsfm.trainer import Trainer  # noqa: E402


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).norm() / (b.double().cpu().norm() + 1e-30))


def oracle_nets(dt=torch.float64):
    d, p = N.DispResNet(18).to(dt), N.PoseResNet(18).to(dt)
    for net in (d, p):
        net.load_state_dict({k: v.to(dt) for k, v in det_weights(net.state_dict()).items()})
        net.train()
    return d, p


def flat_grads(d, p):
    return {"disp." + k: q.grad for k, q in d.named_parameters() if q.grad is not None} | \
           {"pose." + k: q.grad for k, q in p.named_parameters() if q.grad is not None}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    per, H, W = 2, 128, 160                     # 2 frames per rank; 128x160x2 > 10000 pixels: both mean_on_mask branches are live
    tgt, refs, K = synth.triplet(77, per * world, H, W)
    sl = slice(rank * per, (rank + 1) * per)
    args = (tgt[sl].to(dev), [r[sl].to(dev) for r in refs], K[sl].to(dev))
    dt = torch.float64

    def ours(exact, mode="fp32"):
        d, p = models.DispResNet(18, False), models.PoseResNet(18, False)
        for net in (d, p):
            net.load_state_dict(det_weights(net.state_dict()))
        tr = Trainer(d.to(dev).train(), p.to(dev).train(), lr=1e-4, with_auto_mask=1, distributed=True, conv_mode=mode,
                     exact_global_masks=exact)
        return tr, d, p

    # ---- default DDP semantics: mean of the per-shard oracle gradients -----------------------------------------
    tr, d, p = ours(False)
    losses = tr.step(*args)
    got = {"disp." + k: q.grad.clone() for k, q in d.named_parameters()} | {"pose." + k: q.grad.clone() for k, q in p.named_parameters()}
    def shard_mean(dtype):
        acc = None
        for r in range(world):
            od, op = oracle_nets(dtype)
            s = slice(r * per, (r + 1) * per)
            OS.train_step(od, op, OS.make_optimizer(od, op, lr=1e-4), tgt[s].to(dtype), [x[s].to(dtype) for x in refs], K[s].to(dtype),
                          num_scales=1, with_ssim=1, with_mask=1, with_auto_mask=1)
            g = flat_grads(od, op)
            acc = g if acc is None else {k: acc[k] + g[k] for k in g}
        return {k: v / world for k, v in acc.items()}
    want, want32 = shard_mean(dt), shard_mean(torch.float32)
    errs = sorted(rel(got[k], want[k]) for k in want)
    yard = sorted(rel(want32[k], want[k]) for k in want)          # the fp32 CPU oracle's own error on the same quantity
    print("rank %d ddp: all-reduced gradients vs mean of per-shard fp64 oracle gradients: median %.2e worst %.2e (fp32 CPU oracle's own: "
          "median %.2e worst %.2e)" % (rank, errs[len(errs) // 2], errs[-1], yard[len(yard) // 2], yard[-1]), flush=True)
    # a wrong reduction (sum instead of mean, a missing arena, stale replicas) would show up as an error of order 1
    assert errs[len(errs) // 2] < max(4 * yard[len(yard) // 2], 3e-2) and errs[-1] < max(4 * yard[-1], 1e-1), errs[-5:]
    # replicas identical after the update
    for net in (d, p):
        mine = net.flat_params().clone()
        other = mine.clone()
        dist.broadcast(other, 0)
        assert torch.equal(mine, other), "replicas diverged"
    if rank == 0:
        print("DDP_CHECK_OK ddp", flush=True)

    # ---- exact-global masks: the reference's DataParallel semantics ----------------------------------------------
    tr, d, p = ours(True)
    losses = [float(v) for v in tr.step(*args)]
    got = {"disp." + k: q.grad.clone() for k, q in d.named_parameters()} | {"pose." + k: q.grad.clone() for k, q in p.named_parameters()}
    od, op = oracle_nets(dt)
    tds, rds, pss, pis = [], [[] for _ in refs], [[] for _ in refs], [[] for _ in refs]
    for r in range(world):                      # per-GPU-slice forward (BatchNorm statistics per slice), like DataParallel
        s = slice(r * per, (r + 1) * per)
        t64, r64 = tgt[s].to(dt), [x[s].to(dt) for x in refs]
        tds.append([1 / o for o in od(t64)])
        for i, x in enumerate(r64):
            rds[i].append([1 / o for o in od(x)])
            pss[i].append(op(t64, x))
            pis[i].append(op(x, t64))
    cat = lambda lst: [torch.cat([e[s] for e in lst], 0) for s in range(len(lst[0]))]  # noqa: E731
    td = cat(tds)
    rd = [cat(x) for x in rds]
    ps, pi = [torch.cat(x, 0) for x in pss], [torch.cat(x, 0) for x in pis]
    t64, r64, K64 = tgt.to(dt), [x.to(dt) for x in refs], K.to(dt)
    l1, l3 = OL.compute_photo_and_geometry_loss(t64, r64, K64, td, rd, ps, pi, 1, 1, 1, 1, "zeros")
    l2 = OL.compute_smooth_loss(td, t64, rd, r64)
    (l1 + 0.1 * l2 + 0.5 * l3).backward()
    want = flat_grads(od, op)
    # photo / geometry losses are GLOBAL (identical on every rank and equal to the full-batch oracle); smoothness is the rank's own
    assert abs(losses[1] - float(l1)) <= 1e-4 * abs(float(l1)) and abs(losses[3] - float(l3)) <= 1e-4 * abs(float(l3)), (losses, float(l1), float(l3))
    errs = sorted(rel(got[k], want[k]) for k in want)
    print("rank %d exact-global: gradients vs DataParallel-semantics fp64 oracle: median %.2e worst %.2e" % (rank, errs[len(errs) // 2], errs[-1]), flush=True)
    assert errs[len(errs) // 2] < 3e-2 and errs[-1] < 1e-1, errs[-5:]
    if rank == 0:
        print("DDP_CHECK_OK exact", flush=True)

    # ---- the data-parallel step as one CUDA graph (NCCL all-reduce captured on the side stream) -----------------------
    eager, de, _ = ours(False, "tf32x3")
    graphed, dg, _ = ours(False, "tf32x3")
    graphed.capture(*args, allow_distributed=True)
    for it in range(3):
        a = [float(v) for v in eager.step(*args)]
        b = [float(v) for v in graphed.step(*args)]
        for x, y in zip(a, b):
            assert abs(x - y) <= (2e-4 if it == 0 else 5e-3) * abs(x) + 1e-6, (it, a, b)
    torch.cuda.synchronize()
    assert float((dg.flat_params() - de.flat_params()).abs().max()) <= 6.1e-4
    mine = dg.flat_params().clone()
    other = mine.clone()
    dist.broadcast(other, 0)
    assert torch.equal(mine, other), "graphed replicas diverged"
    if rank == 0:
        print("DDP_CHECK_OK graph", flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)          # (a process group must not be torn down while captured graphs still hold its NCCL kernels: it hangs)


if __name__ == "__main__":
    main()
