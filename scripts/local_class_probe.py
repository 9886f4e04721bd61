import sys, time, numpy as np, ctypes as C
sys.path.insert(0,'.')
from __graft_entry__ import load_package
fp = load_package()
from fast_pauli_b200.synth import random_strings
ctx = fp.default_context()
n = 31
rng = np.random.default_rng(5)
psi = ctx.uniform((1<<n,), np.complex128, seed=18)
out = ctx.empty((1<<n,), np.complex128)
for S in (1, 2, 4, 8):
    strings = random_strings(rng, n, S)
    h = rng.uniform(-1,1,S)+1j*rng.uniform(-1,1,S)
    op = fp.PauliOp(h, strings, ctx=ctx)
    plan = op._plan(np.complex128)
    def call():
        fp._check(fp.lib.fp_op_apply(ctx._h, plan, C.c_void_p(out.ptr), C.c_void_p(psi.ptr), C.c_size_t(1<<n), C.c_size_t(1), 0))
    call(); ctx.sync()
    l0 = ctx.launch_count
    t0=time.perf_counter(); call(); ctx.sync(); t1=time.perf_counter()
    print(f"fused PauliOp.apply n=31 B=1 S={S}: {1e3*(t1-t0):.1f} ms, launches {ctx.launch_count-l0}")
    codes,_ = fp._encode(strings)
    t0=time.perf_counter()
    for s in range(S):
        c = np.array([h[s]])
        fp._check(fp.lib.fp_string_apply(ctx._h, fp.FP_C128, n, C.c_void_p(codes[s].ctypes.data), C.c_void_p(c.ctypes.data), C.c_void_p(out.ptr), C.c_void_p(psi.ptr), C.c_size_t(1<<n), C.c_size_t(1), 1))
    ctx.sync(); t1=time.perf_counter()
    print(f"   per-string streaming x{S}: {1e3*(t1-t0):.1f} ms")
