"""Host run of the vectorised sampler-step kernel bodies (fastgen_b200/csrc/sampler_vec.cuh): the same source the
__global__ wrappers of sampler.cu call is compiled with g++ (tests/host_emu/sampler_emu.cpp) and executed for every
thread id of an emulated launch; outputs must equal the reference's formulas evaluated by torch on the CPU BIT FOR BIT,
every vector access must be naturally aligned and inside its buffer, and every output byte written exactly once.

This checks what cannot be seen without a GPU -- index arithmetic, grid-stride tails, alignment -- for kernels that are
opt-in (FASTGEN_B200_SAMPLER_VEC=1) until they have run on hardware.  The emulation library is test infrastructure
(tests/host_emu/build/, git-ignored); the product has no CPU path."""
import ctypes as C
import os
import subprocess

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emu", "sampler_emu.cpp")
OUT = os.path.join(HERE, "host_emu", "build", "libsampler_emu.so")


class EdmStepV(C.Structure):
    _fields_ = [("f", C.c_void_p), ("x_t", C.c_void_p), ("eps", C.c_void_p), ("n", C.c_int), ("c", C.c_int), ("hw", C.c_int),
                ("c_skip", C.c_void_p), ("c_out", C.c_void_p), ("ode", C.c_int), ("alpha_cur", C.c_void_p),
                ("sigma_cur", C.c_void_p), ("alpha_next", C.c_void_p), ("sigma_next", C.c_void_p), ("x0", C.c_void_p),
                ("x_next", C.c_void_p)]


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-Werror", "-ffp-contract=off", "-fPIC", "-shared", "-I", cuda_inc,
                    SRC, "-o", OUT], check=True)
    lib = C.CDLL(OUT)
    lib.emu_message.restype = C.c_char_p
    lib.emu_coverage.restype = C.c_double
    lib.emu_region.argtypes = [C.c_void_p, C.c_int64, C.c_int]
    lib.emu_edm_step.argtypes = [C.POINTER(EdmStepV), C.c_int, C.c_int64]
    lib.emu_coef.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                             C.c_void_p, C.c_void_p, C.c_int64]
    return lib


def _reg(lib, t, writable=0):
    if t is not None:
        lib.emu_region(t.data_ptr(), t.numel() * t.element_size(), writable)


def _ptr(t):
    return t.data_ptr() if t is not None else None


DT = {torch.float32: 0, torch.bfloat16: 1, torch.float64: 2}


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("mode", ["last", "last_noprecond", "sde", "ode"])
@pytest.mark.parametrize("shape,threads", [((5, 3, 16, 16), 7), ((5, 3, 16, 16), 4096), ((2, 4, 8, 8), 33), ((3, 1, 2, 2), 2),
                                           ((2, 3, 6, 10), 256), ((1, 3, 64, 64), 1000)])
def test_edm_step_vector_body_is_bit_exact_on_the_host(emu, dtype, mode, shape, threads):
    torch.manual_seed(5)
    n, c, h, w = shape
    sd = 0.5
    x_t = (torch.randn(n, c, h, w) * 3).to(dtype)
    eps = torch.randn(n, c, h, w).to(dtype)
    f = torch.randn(n, h, w, 4)
    t = torch.rand(n, dtype=torch.float64) * 80 + 0.01
    t_next = torch.rand(n, dtype=torch.float64) * 2 + 0.002
    alpha_next = torch.rand(n, dtype=torch.float64) * 0.5 + 0.5
    alpha_cur = torch.rand(n, dtype=torch.float64) * 0.5 + 0.5
    c_skip = sd**2 / (t**2 + sd**2)
    c_out = t * sd / (t**2 + sd**2).sqrt()
    bc = lambda v: v.reshape(n, 1, 1, 1)  # noqa: E731
    F_ = f[..., :c].permute(0, 3, 1, 2).to(dtype)
    precond = mode != "last_noprecond"
    x0_ref = bc(c_skip.to(dtype)) * x_t + bc(c_out.to(dtype)) * F_ if precond else F_.contiguous()
    x0 = torch.full_like(x_t, float("nan"))
    x_next = torch.full_like(x_t, float("nan")) if mode in ("sde", "ode") else None
    p = EdmStepV()
    p.f, p.x_t, p.n, p.c, p.hw = f.data_ptr(), x_t.data_ptr(), n, c, h * w
    p.c_skip, p.c_out = (_ptr(c_skip), _ptr(c_out)) if precond else (None, None)
    p.x0, p.x_next = x0.data_ptr(), _ptr(x_next)
    if mode == "sde":
        p.eps, p.alpha_next, p.sigma_next = eps.data_ptr(), alpha_next.data_ptr(), t_next.data_ptr()
        e = eps
    elif mode == "ode":
        p.ode, p.alpha_next, p.sigma_next = 1, alpha_next.data_ptr(), t_next.data_ptr()
        p.alpha_cur, p.sigma_cur = alpha_cur.data_ptr(), t.data_ptr()
        e = ((x_t.double() - x0_ref.double() * bc(alpha_cur)) / bc(t)).to(dtype)
    emu.emu_reset()
    for ten in (f, x_t, eps if mode == "sde" else None):
        _reg(emu, ten)
    _reg(emu, x0, 1)
    _reg(emu, x_next, 1)
    emu.emu_edm_step(C.byref(p), DT[dtype], threads)
    assert emu.emu_errors() == 0, emu.emu_message().decode()
    assert emu.emu_coverage(0) == 1.0 and (x_next is None or emu.emu_coverage(1) == 1.0)
    assert torch.equal(x0, x0_ref)
    if x_next is not None:
        xn_ref = (x0_ref.double() * bc(alpha_next) + e.double() * bc(t_next)).to(dtype)
        assert torch.equal(x_next, xn_ref)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float64])
@pytest.mark.parametrize("threads", [5, 1 << 14])
def test_coef_vector_bodies_are_bit_exact_on_the_host(emu, dtype, threads):
    torch.manual_seed(6)
    B, Cc, T, H, W = 4, 16, 3, 10, 12
    a = torch.randn(B, Cc, T, H, W).to(dtype)
    b = torch.randn_like(a)
    total = a.numel()

    def run(mode, a_, b_, outer, mid, inner, c0, c1):
        out = torch.full_like(a_, float("nan"))
        emu.emu_reset()
        _reg(emu, a_)
        _reg(emu, b_)
        _reg(emu, out, 1)
        emu.emu_coef(mode, DT[dtype], a_.data_ptr(), _ptr(b_), total, outer, mid, inner, c0.data_ptr(), _ptr(c1), out.data_ptr(),
                     threads)
        assert emu.emu_errors() == 0, emu.emu_message().decode()
        assert emu.emu_coverage(0) == 1.0
        return out

    # per-sample coefficients [B]
    ca = torch.rand(B, dtype=torch.float64)
    cb = torch.rand(B, dtype=torch.float64) + 0.1
    bc = lambda v: v.reshape(B, 1, 1, 1, 1)  # noqa: E731
    per = total // B
    assert torch.equal(run(0, a, b, per, 1, per, ca, cb), (a.double() * bc(ca) + b.double() * bc(cb)).to(dtype))
    assert torch.equal(run(0, a, None, per, 1, per, ca, None), (a.double() * bc(ca)).to(dtype))
    assert torch.equal(run(1, a, b, per, 1, per, ca, cb), ((a.double() - b.double() * bc(ca)) / bc(cb)).to(dtype))
    # per-frame coefficients [B, T] of a [B, C, T, H, W] latent (causvid.py:55-56): outer = C*T*H*W, mid = T, inner = H*W
    fa = torch.rand(B, T, dtype=torch.float64)
    fb = torch.rand(B, T, dtype=torch.float64) + 0.1
    bf = lambda v: v.reshape(B, 1, T, 1, 1)  # noqa: E731
    assert torch.equal(run(0, a, b, per, T, H * W, fa, fb), (a.double() * bf(fa) + b.double() * bf(fb)).to(dtype))
    assert torch.equal(run(1, a, b, per, T, H * W, fa, fb), ((a.double() - b.double() * bf(fa)) / bf(fb)).to(dtype))


def test_the_access_checker_catches_misaligned_and_out_of_range_accesses(emu):
    """the emulation is only worth something if it fails when the kernel would fault"""
    a = torch.randn(2, 64)
    out = torch.empty_like(a)
    c0 = torch.ones(2, dtype=torch.float64)
    emu.emu_reset()
    _reg(emu, a)
    _reg(emu, out, 1)
    emu.emu_coef(0, 0, a.data_ptr() + 4, None, 124, 62, 1, 62, c0.data_ptr(), None, out.data_ptr(), 8)   # shifted by one float
    assert emu.emu_errors() > 0 and b"misaligned" in emu.emu_message()
    emu.emu_reset()
    _reg(emu, a)
    _reg(emu, out[:1], 1)                                                                           # output buffer too small
    emu.emu_coef(0, 0, a.data_ptr(), None, 128, 64, 1, 64, c0.data_ptr(), None, out.data_ptr(), 8)
    assert emu.emu_errors() > 0 and b"outside" in emu.emu_message()
