"""Does a solve with the persisting-L2 window leave the device as it found it?  7-point convection-diffusion 256^3 BiCGSTAB
(16.8 M rows: the per-GPU share of configs[4] on 8 GPUs, no window) timed before and after a 7-point 128^3 CG solve (window on),
plus the carve-out limit as the runtime reports it."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from liblcg_b200 import api


def limit():
    cudart = C.CDLL("libcudart.so.12")
    v = C.c_size_t()
    cudart.cudaDeviceGetLimit(C.byref(v), 6)   # cudaLimitPersistingL2CacheSize
    return v.value


def rate(ctx, kind, g, solver, iters, steps=3):
    torch = ctx.torch
    S = bench.device_system(ctx, kind, g, solver)
    m = torch.zeros(S["n_loc"], dtype=torch.float64, device=ctx.dev)
    para = api.lcg_default_parameters(epsilon=1e-300, max_iterations=iters)
    step = lambda: api.solve(S["op"], bench.SOLVER_ID[solver], m, S["b"], param=para, device=True, stream=ctx.stream)
    for _ in range(2):
        m.zero_(); step()
    ms, _ = ctx.timed(step, steps, prepare=lambda: m.zero_())
    bench.close_system(ctx, S)
    return steps * iters / (ms * 1e-3)


ctx = bench.Ctx()
print("carve-out at start:", limit())
before = rate(ctx, "7pt_cd", 256, "BICGSTAB", 60)
print(f"bicgstab 7pt_cd 256^3 before: {before:.1f} it/s, carve-out {limit()}")
mid = rate(ctx, "7pt", 128, "CG", 200)
print(f"cg 7pt 128^3 (window on): {mid:.1f} it/s, carve-out afterwards {limit()}")
after = rate(ctx, "7pt_cd", 256, "BICGSTAB", 60)
print(f"bicgstab 7pt_cd 256^3 after: {after:.1f} it/s ({after / before:.3f} of before), carve-out {limit()}")
