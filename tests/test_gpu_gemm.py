"""tcgen05 GEMM kernel (csrc/gemm_tc.cuh) against a float64 reference on the same bf16-rounded operands."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _gelu(x):
    return 0.5 * x * (1.0 + np.tanh(np.sqrt(2.0 / np.pi) * (x + 0.044715 * x ** 3)))


@pytest.fixture(scope="module", params=["bf16", "f16"])
def eng(request):
    """One handle per operand-format build of the library; operands are generated in that build's 16-bit format."""
    from gpu_util import engine, make_model
    m, _ = make_model(1, precision=request.param)
    e = engine(m)
    e.op_dtype = torch.bfloat16 if request.param == "bf16" else torch.float16
    e.keep_model = m
    return e


@pytest.mark.parametrize("bn,M,N,K", [
    (64, 128, 64, 64), (64, 1000, 320, 256), (128, 128, 128, 64), (128, 4096, 256, 512),
    (128, 333, 128, 128), (256, 256, 256, 128), (256, 20000, 512, 256), (128, 16384, 256, 256),
])
def test_gemm_plain(eng, bn, M, N, K):
    from gpu_util import debug_gemm
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g).to(eng.op_dtype).cuda()
    w = (torch.randn(N, K, generator=g) / np.sqrt(K)).to(eng.op_dtype).cuda()
    out32, _ = debug_gemm(eng, bn, a, w, 16)
    ref = a.double().cpu().numpy() @ w.double().cpu().numpy().T
    err = np.abs(out32.cpu().numpy() - ref).max()
    assert err < 2e-4 * np.sqrt(K), f"max abs err {err}"       # fp32 accumulation of exact bf16 products


def test_gemm_fused_epilogue(eng):
    from gpu_util import debug_gemm
    M, N, K = 1500, 256, 128
    g = torch.Generator(device="cpu").manual_seed(5)
    a = torch.randn(M, K, generator=g).to(eng.op_dtype).cuda()
    w = (torch.randn(N, K, generator=g) / np.sqrt(K)).to(eng.op_dtype).cuda()
    bias = torch.randn(N, generator=g).cuda()
    gamma = torch.rand(N, generator=g).cuda() + 0.5
    resid = torch.randn(M, N, generator=g).cuda()
    acc = a.double().cpu().numpy() @ w.double().cpu().numpy().T + bias.double().cpu().numpy()
    # bias + gelu -> bf16
    _, o16 = debug_gemm(eng, 128, a, w, 1 | 2 | 32, bias=bias, want32=False, want16=True)
    ref = _gelu(acc)
    assert np.abs(o16.float().cpu().numpy() - ref).max() < 2e-2          # bf16 output rounding (|x| < 6)
    # bias + gamma + residual -> fp32 (the pw2 / ffn2 / out-proj epilogue)
    o32, _ = debug_gemm(eng, 256, a, w, 1 | 4 | 8 | 16, bias=bias, gamma=gamma, resid=resid)
    ref = resid.double().cpu().numpy() + gamma.double().cpu().numpy() * acc
    assert np.abs(o32.cpu().numpy() - ref).max() < 1e-3


def test_gemm_strided_a(eng):
    """A operand read through a row stride (K = 64 slice of a [M, 320] buffer, as the kv projection does)."""
    from gpu_util import debug_gemm
    M, N, K = 700, 128, 64
    g = torch.Generator(device="cpu").manual_seed(9)
    full = torch.randn(M, 320, generator=g).to(eng.op_dtype).cuda()
    a = full[:, 256:]
    w = (torch.randn(N, K, generator=g) / 8).to(eng.op_dtype).cuda()
    o32, _ = debug_gemm(eng, 128, a, w, 16)
    ref = a.double().cpu().numpy() @ w.double().cpu().numpy().T
    assert np.abs(o32.cpu().numpy() - ref).max() < 2e-3


@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (512, 128, 256), (1024, 256, 256), (256, 256, 512)])
def test_cta_pair_gemm(M, N, K):
    """tcgen05 cta_group::2 (gemm_pair.cuh): a cluster of two CTAs computes 256 rows, each CTA holding half of W."""
    import ctypes as C
    from audio_to_midi_b200 import _lib
    from gpu_util import engine, make_model
    m, _ = make_model(1, precision="bf16")
    eng = engine(m)
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A_ = torch.randn(M, K, generator=g).to(torch.bfloat16).cuda()
    W = (torch.randn(N, K, generator=g) * 0.1).to(torch.bfloat16).cuda()
    out = torch.full((M, N), float("nan"), dtype=torch.float32, device="cuda")
    f = eng.L.a2m_debug_gemm_pair
    f.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    f.restype = C.c_int
    rc = f(eng.h, M, N, K, A_.data_ptr(), W.data_ptr(), out.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(eng.h, rc, "a2m_debug_gemm_pair")
    torch.cuda.synchronize()
    ref = A_.double() @ W.double().T
    err = (out.double() - ref).abs().max().item()
    assert err < 1e-3 * np.sqrt(K), err
