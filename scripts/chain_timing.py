"""Stage timing of the fused backward chain: builds an instrumented copy of the library (-DGCNN_CHAIN_TIMING), runs
train steps on the bench workload and prints the clock64() deltas thread 0 of CTA 0 recorded in its first tiles.
Usage (GPU box): python scripts/chain_timing.py"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "gcnn_cut_selector_b200")
OUT = os.path.join(PKG, "build", "libgcnn_b200_timing.so")
os.environ["GCNN_LIB"] = OUT  # read when gcnn_cut_selector_b200._lib is first imported


def build():
    from gcnn_cut_selector_b200 import build as b
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    flags = [f for f in b.NVCC_FLAGS if not f.startswith("--use_fast_math")]
    srcs = [os.path.join(b.CSRC, s) for s in b.SOURCES]
    subprocess.run(["/usr/local/cuda/bin/nvcc", *flags, "-DGCNN_CHAIN_TIMING", "-shared", "-o", OUT, *srcs, "-cudart", "static"],
                   check=True)


if __name__ == "__main__":
    if "--build" in sys.argv or not os.path.exists(OUT):
        build()
    if "--build-only" in sys.argv:
        sys.exit(0)
    import torch
    import bench
    from gcnn_cut_selector_b200 import GCNN, _lib, batching
    dev = torch.device("cuda:0")
    model = GCNN(device=dev, seed=0)
    model.check_indices = False
    batches = bench.make_batches(2, 32, seed0=0)
    inputs = [model.prepare_inputs(batching.model_inputs(b, per_sample_counts=True)) for b in batches]
    targets = [torch.from_numpy(b[10]).to(dev) for b in batches]
    for i in range(4):
        model.loss_and_grads(inputs[i % 2], targets[i % 2])
        model.apply_gradients(1e-4)
    torch.cuda.synchronize()
    lib = _lib.load()
    lib.gcnn_debug_chain_timing.restype = C.c_int
    buf = (C.c_longlong * 64)()
    n = lib.gcnn_debug_chain_timing(buf, 64)
    ts = [buf[i] for i in range(n)]
    print("tc_conv_backward_kernel, thread 0 of CTA 0, last launch of the step (convolution 0: one 128-node tile per CTA).\n"
          "Marks in code order: 0 kernel entry | 1 TMEM + barriers ready | 2 previous grid done (griddepcontrol.wait) |\n"
          "3 prologue done | 4 tile start | S3: 5 tiles stored, 6 published, 7 prefetch issued, 8 input-gradient MMAs done,\n"
          "9 epilogue done | S2: 10 weight-gradient MMAs done, 11 stored, 12 published, 13 prefetch issued, 14 MMAs done,\n"
          "15 epilogue done | S1: 16 weight-gradient MMAs done, 17 stored, 18 published, 19 prefetch issued, 20 MMAs done,\n"
          "21 weight-gradient MMAs done (epilogue part 1), 22 stored | S0: 23 published, 24 prefetch issued, 25 MMAs done,\n"
          "26 epilogue done (tiles done) | 27 accumulators and bias sums drained.  1.965 GHz: 1,000 cycles = 0.51 us.")
    print("marks:", n)
    for i in range(1, n):
        print(f"  {i:2d}: +{ts[i] - ts[i - 1]:7d} cycles  (t={ts[i] - ts[0]})")
