"""-m gpu: op-level parity of the CUDA kernels (through the C ABI) against torch fp32 / the oracle."""
import pytest

import gpu_checks as G

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(128, 96, 96), (1000, 288, 96), (4096, 384, 96), (512, 96, 384), (300, 768, 768),
                                   (2048, 2304, 768), (640, 527, 4608), (4096, 256, 1024)])
@pytest.mark.parametrize("out_bf16", [False, True])
def test_gemm(M, N, K, out_bf16):
    # operands are bf16-rounded on both sides; fp32 accumulation => only summation-order noise (+ bf16 output rounding)
    r, _ = G.check_gemm(M, N, K, out_bf16)
    assert r < (4e-3 if out_bf16 else 2e-5), r


@pytest.mark.parametrize("M,N,K", [(2048, 2304, 768), (5000, 1536, 384), (65536, 384, 1536), (2304, 128, 384), (4097, 192, 3072),
                                   (300 * 128, 1152, 384)])
@pytest.mark.parametrize("out_bf16", [False, True])
def test_gemm_cta_pair(M, N, K, out_bf16, monkeypatch):
    """The cta_group::2 (CTA-pair, M=256) kernel, forced for every shape incl. M not a multiple of 256."""
    monkeypatch.setenv("ARD_GEMM_PAIR", "1")
    r, _ = G.check_gemm(M, N, K, out_bf16, nres=0 if out_bf16 else 1)
    assert r < (4e-3 if out_bf16 else 2e-5), r


def test_gemm_epilogues():
    assert G.check_gemm(4096, 384, 96, True, act=1)[0] < 4e-3      # exact-erf GELU
    assert G.check_gemm(512, 512, 768, False, act=2)[0] < 2e-5     # ReLU
    assert G.check_gemm(4096, 96, 384, False, nres=2)[0] < 2e-5    # two residual adds (patched block, src/residual.py:95)
    assert G.check_gemm(1024, 192, 384, False, bias=False, nres=1)[0] < 2e-5


@pytest.mark.parametrize("M,resid2", [(128, False), (5000, True), (16384 + 77, True), (148 * 128 * 3 + 77, True), (148 * 128 * 2, False)])
def test_ffn_fused_96(M, resid2):
    r_out, r_branch = G.check_ffn_fused(M, resid2)
    assert r_out < 2e-3 and r_branch < 5e-3, (r_out, r_branch)     # bf16 LN output / fp16 hidden operands, fp32 accumulation


@pytest.mark.parametrize("C,M,resid2,alias", [(192, 128, False, False), (192, 5000, True, True), (192, 148 * 128 * 3 + 77, True, False),
                                             (384, 100, False, True), (384, 5000, True, False), (384, 148 * 128 * 2 + 300, False, True),
                                             (128, 128, False, False), (128, 5000, True, True), (128, 148 * 128 * 3 + 77, True, False),
                                             (256, 100, False, True), (256, 5000, True, False), (256, 148 * 128 * 2 + 300, False, True)])
def test_ffn_fused_wide(C, M, resid2, alias):
    r_out, r_branch = G.check_ffn_fused(M, resid2, Cd=C, alias=alias)
    assert r_out < 2e-3 and r_branch < 5e-3, (C, M, r_out, r_branch)


@pytest.mark.parametrize("M", [128, 5000, 148 * 128 * 3 + 77])
def test_ln_qkv_96(M):
    assert G.check_ln_qkv(M) < 4e-3, M          # bf16 output rounding (2^-9) on top of bf16 operands


@pytest.mark.parametrize("M,N,K", [(128, 384, 96), (4096 + 64, 384, 96), (3000, 768, 192), (20000, 1536, 384), (640, 512, 128)])
def test_gemm_dual_gelu_backward(M, N, K):
    """dh = (g W2) * gelu'(fc1(norm2(x)) + b1) in one kernel (FFN backward): packed-fp16 gelu' (fit error 1.2e-4, fp16 evaluation
    4e-4) on top of the bf16 output rounding."""
    r, _ = G.check_gemm_dual(0, M, N, K)
    assert r < 4e-3, r


@pytest.mark.parametrize("M,N,K,Kvalid", [(4096, 96, 96, 96), (4096 + 64, 48, 96, 40), (3000, 192, 192, 192), (2048, 384, 384, 384),
                                          (1100, 768, 768, 768), (200000, 96, 96, 96)])
def test_gemm_dual_lambda_gradient(M, N, K, Kvalid):
    """ResiDual backward: dlam += colsum(x_proj * dL/d x_scaled) reduced in the kernel (registers -> shared -> one atomic per column
    per CTA), gsc = gcoef * lambda as the bf16 output. N = component count padded to 16, Kvalid = the real one."""
    r_out, r_dl = G.check_gemm_dual(1, M, N, K, Kvalid)
    assert r_out < 4e-3 and r_dl < 1e-4, (r_out, r_dl)


def test_gemm_f16_hidden_chain():
    r_h, r_out = G.check_gemm_f16_chain()
    assert r_h < 1e-3 and r_out < 2e-3, (r_h, r_out)                # fp16 hidden: 10-bit mantissa


@pytest.mark.parametrize("C", [96, 128, 192, 256, 384, 512, 768, 1024, 1536, 2048])
def test_layernorm(C):
    r_bf16, r_f32 = G.check_layernorm(777, C)
    assert r_bf16 < 1e-3 and r_f32 < 4e-3, (r_bf16, r_f32)


@pytest.mark.parametrize("B,R,C,nH,shift", [(2, 64, 96, 4, 0), (2, 64, 96, 4, 4), (2, 32, 192, 8, 4), (3, 16, 384, 16, 4),
                                            (2, 8, 768, 32, 4), (2, 64, 128, 4, 4), (2, 16, 512, 16, 0), (1, 32, 256, 8, 4)])
def test_window_attention(B, R, C, nH, shift):
    r_out, r_attn = G.check_window_attention(B, R, C, nH, shift)
    assert r_out < G.TOL_BF16 and r_attn < 1e-3, (r_out, r_attn)    # probabilities are fp32 softmax of bf16-exact logits


def test_logmel():
    m = G.check_logmel()
    assert m["logmel"] < G.TOL_FP32 and m["logmel_bn"] < 2e-4 and m["logmel_maxabs_dB"] < 5e-3, m


@pytest.mark.parametrize("rows,D,strided,calls", [(1000, 96, False, 1), (5000, 192, False, 2), (777, 768, False, 1), (3000, 4096, True, 2),
                                                  (130, 384, True, 1)])
def test_stats_accumulate(rows, D, strided, calls):
    m = G.check_stats(rows, D, strided, calls=calls)
    assert m["n"] == 0 and m["sum"] < 1e-6 and m["sumsq"] < 1e-5, m


def test_quantize_waveform_bit_exact():
    assert G.check_quantize_waveform() == 0


@pytest.mark.parametrize("B,block,residual", [(1, 0, False), (2, 1, False), (3, 0, True), (2, 1, True), (150, 1, True)])
def test_attention_block_fused(B, block, residual):
    """The window-resident tcgen05 attention block of the 96-channel stage (shifted and unshifted windows, plain and
    ResiDual-folded projection, more tiles than SMs) vs the oracle and vs the unfused kernel chain."""
    if B > 8:
        # the oracle at B=150 takes minutes: compare with the unfused kernel chain only
        import torch
        clap, sd, _ = G.make_encoder("tiny", residual=residual)
        enc = clap.model.audio_branch
        x = (torch.randn(B, 4096, 96, generator=torch.Generator().manual_seed(1)) * 0.8).cuda()
        out = torch.empty_like(x)
        G.L.check(G.L.load().ard_attention_block(enc._handle(), 0, block, G.L.ptr(x), B, G.L.ptr(out), G.L.stream_ptr()))
        _, _, res = enc.layers[0].blocks[block](x)
        torch.cuda.synchronize()
        assert G.rel(out - x, res) < G.TOL_BF16
        return
    m = G.check_attention_block(B, block, residual)
    assert m["out"] < 2e-3 and m["branch"] < G.TOL_BF16 and m["branch_vs_unfused"] < G.TOL_BF16, m
