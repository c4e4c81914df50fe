"""CHECKER (test infrastructure): execute a recorded auto-SPMD bundle of the reference's benchmark
GPT (tests/golden/auto_*.json.gz: traced graph + the plan the unmodified reference's solver chose)
with this backend's lowering and executor on N ranks and compare outputs, every parameter and every
momentum buffer with vanilla PyTorch (the reference's comparator, tests/test_torch/test_spmd.py:
97-113; rtol 1e-4 / atol 1e-5).  CPU: gloo stand-ins for the kernels (tests/gloo_ops.py), e.g.

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/validate_bundle.py \\
      --bundle tests/golden/auto_gpt2medium_s128_mesh8.json.gz --gpt 24,1024,16,8,128

On GPUs (`--device cuda`): the product path (libedb kernels), as tests/mgpu_worker.py does for the
config-1 bundles."""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bundle", required=True)
    ap.add_argument("--gpt", required=True, help="depth,dim,heads,batch,seq the bundle was solved for")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--device", default="cpu", choices=("cpu", "cuda"))
    ap.add_argument("--threads", type=int, default=1)
    ap.add_argument("--heap-gb", type=float, default=8.0)
    ap.add_argument("--dtype", default="fp32", choices=("fp32", "bf16"),
                    help="bf16: parameters and inputs in bf16 (plans are dtype independent); the "
                         "comparison with vanilla bf16 PyTorch is then a relative L2 error <= 5e-2 per tensor (two bf16 runs with different summation orders: the bench parity legs measure 3-8e-2 between vanilla bf16 and fp32)")
    ap.add_argument("--vanilla-ranks", default="", help="e.g. 0: only these ranks hold the vanilla "
                    "model and compare (host memory); default all")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    depth, dim, heads, batch, seq = (int(v) for v in args.gpt.split(","))
    torch.set_num_threads(args.threads)
    from tests.test_auto_bundle_cpu import run_c1_bundle
    if args.device == "cuda":
        local = int(os.environ.get("LOCAL_RANK", rank))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        from easydist_b200 import reshard as ops, runtime
        runtime.init(rank, world, local, heap_bytes=int(args.heap_gb * (1 << 30)))
        native = True
    else:
        dist.init_process_group("gloo")
        from tests import gloo_ops as ops
        ops.init_groups(np.arange(world).reshape((world,)))
        native = False
    t0 = time.time()
    ok, msg, hist = run_c1_bundle(rank, world, ops, native, args.device, steps=args.steps,
                                  tag=str(world), bundle_file=args.bundle, gpt=(depth, dim, heads),
                                  batch=batch, seq=seq,
                                  **(dict(dtype=torch.bfloat16, rtol=None, atol=5e-2)
                                     if args.dtype == "bf16" else {}),
                                  vanilla_ranks=[int(v) for v in args.vanilla_ranks.split(",")]
                                  if args.vanilla_ranks else None)
    flag = torch.tensor([0.0 if ok else 1.0], device=args.device)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"VALIDATE_BUNDLE ok={flag.item() == 0.0} world={world} gpt={args.gpt} dtype={args.dtype} steps={args.steps} "
              f"localize={os.environ.get('EDB_LOCALIZE_OPT', '1' if native else '0')} comm={hist} "
              f"t={time.time() - t0:.0f}s {msg}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if flag.item() != 0.0:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
