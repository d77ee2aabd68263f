"""Parity checks of the CUDA path against the CPU oracle (oracle/htsat_oracle.py) and the committed golden vectors.

Shared by tests/test_gpu_*.py (pytest -m gpu) and tools/gpu_debug.py (prints every metric without stopping).
Every check returns a dict of named relative errors; `TOL` holds the stated tolerance for each.

Tolerances (BASELINE.json north_star): bf16 tensor-core path rel. err <= 1e-2 on embeddings / per-layer outputs;
fp32 CUDA-core pieces (front end, LayerNorm statistics, heads) <= 1e-4.  rel err = ||a-b||_2 / ||b||_2.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from audio_residual_b200 import lib as L  # noqa: E402
from audio_residual_b200 import weights as W  # noqa: E402
from oracle import htsat_oracle as O  # noqa: E402

TOL_BF16 = 1e-2
TOL_FP32 = 1e-4
GOLDEN = os.path.join(ROOT, "tests", "golden")


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def golden_sample(t, n=4096):
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    while step > 1 and (step % 2 == 0 or step % 3 == 0):
        step -= 1
    return f[::step][:n].to(torch.float32).cpu()


def bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


# ---------------------------------------------------------------------------------------------------- op level
def check_gemm(M, N, K, out_bf16, act=0, bias=True, nres=0, seed=0):
    lib = L.load()
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(M, K, generator=g)
    Wt = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g) if bias else None
    r1 = torch.randn(M, N, generator=g) if nres >= 1 else None
    r2 = torch.randn(M, N, generator=g) if nres >= 2 else None
    ref = bf16r(A) @ bf16r(Wt).t()
    if b is not None:
        ref = ref + b
    if act == 1:
        ref = torch.nn.functional.gelu(ref)
    elif act == 2:
        ref = torch.relu(ref)
    if r1 is not None:
        ref = ref + r1
    if r2 is not None:
        ref = ref + r2
    dev = "cuda"
    Ad, Wd = A.to(dev, torch.bfloat16).contiguous(), Wt.to(dev, torch.bfloat16).contiguous()
    ldo = N if (N % 8 == 0) else ((N + 7) // 8) * 8
    out = torch.zeros(M, ldo, device=dev, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    bd = b.to(dev) if b is not None else None
    r1d = r1.to(dev).contiguous() if r1 is not None else None
    r2d = r2.to(dev).contiguous() if r2 is not None else None
    L.check(lib.ard_gemm_bf16(L.ptr(Ad), K, L.ptr(Wd), K, L.ptr(out), ldo, int(out_bf16), M, N, K, L.ptr(bd), act,
                              L.ptr(r1d), N, L.ptr(r2d), N, L.stream_ptr()))
    torch.cuda.synchronize()
    got = out[:, :N].float().cpu()
    return rel(got, ref), (got - ref).abs().max().item()


def check_gemm_dual(mode, M, N, K, Kvalid=None, seed=0):
    """ard_gemm_dual vs torch fp32 on bf16-rounded operands. mode 0: out = (A1 W1^T) * gelu'(A2 W2^T + b); mode 1: dlam += colsum((A1 W1^T + c0)
    * (A2 W2^T)), out = (A2 W2^T) * lam. Returns (rel err of out, rel err of dlam or 0)."""
    lib = L.load()
    g = torch.Generator().manual_seed(seed)
    A1, A2 = torch.randn(M, K, generator=g), torch.randn(M, K, generator=g)
    W1, W2 = torch.randn(N, K, generator=g) / K ** 0.5, 1.5 * torch.randn(N, K, generator=g) / K ** 0.5
    v1, v2 = 0.3 * torch.randn(N, generator=g), 1 + 0.2 * torch.randn(N, generator=g)
    Kvalid = N if Kvalid is None else Kvalid
    acc1, acc2 = bf16r(A1) @ bf16r(W1).t(), bf16r(A2) @ bf16r(W2).t()
    d = lambda t: t.cuda().contiguous()   # noqa: E731
    A1d, A2d, W1d, W2d = (d(t).to(torch.bfloat16) for t in (A1, A2, W1, W2))
    v1d, v2d = d(v1), d(v2)
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    dlam = torch.full((Kvalid,), 0.5, device="cuda")   # accumulated into
    if mode == 0:
        h = (acc2 + v1).double()
        gd = 0.5 * (1 + torch.erf(h / 2 ** 0.5)) + h * torch.exp(-h * h / 2) / (2 * torch.pi) ** 0.5
        ref, ref_dl = acc1 * gd.float(), None
    else:
        ref = acc2 * v2
        ref_dl = ((acc1 + v1).double() * acc2.double()).sum(0)[:Kvalid].float() + 0.5
    L.check(lib.ard_gemm_dual(mode, L.ptr(A1d), K, L.ptr(W1d), K, L.ptr(A2d), K, L.ptr(W2d), K, L.ptr(out), N, M, N, K, L.ptr(v1d), L.ptr(v2d),
                              L.ptr(dlam) if mode == 1 else None, Kvalid, L.stream_ptr()))
    torch.cuda.synchronize()
    r_out = rel(out.float().cpu(), ref)
    r_dl = rel(dlam.cpu(), ref_dl) if mode == 1 else 0.0
    return r_out, r_dl


def check_ffn_fused(M=5000, resid2=True, seed=0, Cd=96, alias=False):
    """ard_ffn_fused_96 / ard_ffn_fused_wide (LN + fc1 + GELU + fc2 + residuals in one kernel) vs torch fp32 on bf16/fp16-rounded
    weights. `alias`: the output overwrites x (how the forward schedule calls it)."""
    lib = L.load()
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(M, Cd, generator=g) * 1.5 + 0.3
    r2 = torch.randn(M, Cd, generator=g) if resid2 else None
    gm, bt = 1 + 0.1 * torch.randn(Cd, generator=g), 0.1 * torch.randn(Cd, generator=g)
    w1 = torch.randn(4 * Cd, Cd, generator=g) / Cd ** 0.5
    w2 = torch.randn(Cd, 4 * Cd, generator=g) / (4 * Cd) ** 0.5
    b1, b2 = 0.1 * torch.randn(4 * Cd, generator=g), 0.1 * torch.randn(Cd, generator=g)
    ln = torch.nn.functional.layer_norm(x, (Cd,), gm, bt, 1e-5)
    hdn = torch.nn.functional.gelu(ln @ bf16r(w1).t() + b1)
    branch = hdn @ w2.half().float().t() + b2
    ref = x + branch + (r2 if resid2 else 0)
    d = lambda t: t.cuda().contiguous()
    xd, r2d, gd, btd = d(x), (d(r2) if resid2 else None), d(gm), d(bt)
    w1d, w2d, b1d, b2d = d(w1).to(torch.bfloat16), d(w2).to(torch.float16), d(b1), d(b2)
    out = xd if alias else torch.empty_like(xd)
    if Cd == 96:
        L.check(lib.ard_ffn_fused_96(L.ptr(xd), L.ptr(r2d), L.ptr(out), M, L.ptr(gd), L.ptr(btd), L.ptr(w1d), L.ptr(b1d), L.ptr(w2d), L.ptr(b2d),
                                     L.stream_ptr()))
    else:
        b1h = (0.5 * b1d).contiguous()
        L.check(lib.ard_ffn_fused_wide(L.ptr(xd), L.ptr(r2d), L.ptr(out), M, Cd, L.ptr(gd), L.ptr(btd), L.ptr(w1d), L.ptr(b1h), L.ptr(w2d),
                                       L.ptr(b2d), L.stream_ptr()))
    torch.cuda.synchronize()
    got_branch = out.cpu() - x - (r2 if resid2 else 0)
    return rel(out.cpu(), ref), rel(got_branch, branch)


def check_ln_qkv(M=5000, seed=0):
    """ard_ln_qkv_96 (norm1 + qkv projection in one kernel) vs torch fp32 LayerNorm + Linear on the bf16-rounded weight."""
    lib = L.load()
    g = torch.Generator().manual_seed(seed)
    Cd = 96
    x = torch.randn(M, Cd, generator=g) * 1.5 + 0.3
    gm, bt = 1 + 0.1 * torch.randn(Cd, generator=g), 0.1 * torch.randn(Cd, generator=g)
    w = torch.randn(3 * Cd, Cd, generator=g) / Cd ** 0.5
    b = 0.1 * torch.randn(3 * Cd, generator=g)
    ref = bf16r(torch.nn.functional.layer_norm(x, (Cd,), gm, bt, 1e-5)) @ bf16r(w).t() + b
    d = lambda t: t.cuda().contiguous()
    xd, gd, btd, wd, bd = d(x), d(gm), d(bt), d(w).to(torch.bfloat16), d(b)
    out = torch.empty(M, 3 * Cd, device="cuda", dtype=torch.bfloat16)
    L.check(lib.ard_ln_qkv_96(L.ptr(xd), L.ptr(gd), L.ptr(btd), L.ptr(wd), L.ptr(bd), L.ptr(out), M, L.stream_ptr()))
    torch.cuda.synchronize()
    return rel(out.float().cpu(), ref)


def check_gemm_f16_chain(M=3000, Cd=192, seed=0):
    """fc1 with the packed-fp16 GELU epilogue (ARD_ACT_GELU_F16) feeding the fp16 fc2 GEMM with a TMA-fetched residual."""
    lib = L.load()
    g = torch.Generator().manual_seed(seed)
    xn = torch.randn(M, Cd, generator=g)
    w1 = torch.randn(4 * Cd, Cd, generator=g) / Cd ** 0.5
    w2 = torch.randn(Cd, 4 * Cd, generator=g) / (4 * Cd) ** 0.5
    b1, b2 = 0.1 * torch.randn(4 * Cd, generator=g), 0.1 * torch.randn(Cd, generator=g)
    res = torch.randn(M, Cd, generator=g)
    hdn = torch.nn.functional.gelu(bf16r(xn) @ bf16r(w1).t() + b1)
    ref = hdn @ w2.half().float().t() + b2 + res
    d = lambda t: t.cuda().contiguous()
    xd, w1d, w2d = d(xn).to(torch.bfloat16), d(w1).to(torch.bfloat16), d(w2).to(torch.float16)
    b1d, b2d, rd = d(b1), d(b2), d(res)
    hb = torch.empty(M, 4 * Cd, device="cuda", dtype=torch.float16)
    out = torch.empty(M, Cd, device="cuda", dtype=torch.float32)
    st = L.stream_ptr()
    L.check(lib.ard_gemm_bf16(L.ptr(xd), Cd, L.ptr(w1d), Cd, L.ptr(hb), 4 * Cd, 1, M, 4 * Cd, Cd, L.ptr(b1d), L.ACT_GELU_F16, None, 0, None, 0, st))
    L.check(lib.ard_gemm_f16(L.ptr(hb), 4 * Cd, L.ptr(w2d), 4 * Cd, L.ptr(out), Cd, 0, M, Cd, 4 * Cd, L.ptr(b2d), 0, L.ptr(rd), Cd, None, 0, st))
    torch.cuda.synchronize()
    return rel(hb.float().cpu(), hdn), rel(out.cpu(), ref)


def check_layernorm(rows, Cdim, seed=0):
    lib = L.load()
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(rows, Cdim, generator=g) * 2 + 0.5
    gm = 1 + 0.1 * torch.randn(Cdim, generator=g)
    bt = 0.1 * torch.randn(Cdim, generator=g)
    ref = torch.nn.functional.layer_norm(x, (Cdim,), gm, bt, 1e-5)
    out = torch.empty(rows, Cdim, device="cuda", dtype=torch.bfloat16)
    xd, gd, bd = x.cuda(), gm.cuda(), bt.cuda()   # keep the device copies alive across the async launch
    L.check(lib.ard_layernorm_bf16(L.ptr(xd), L.ptr(gd), L.ptr(bd), L.ptr(out), rows, Cdim, L.stream_ptr()))
    torch.cuda.synchronize()
    return rel(out.float().cpu(), bf16r(ref)), rel(out.float().cpu(), ref)


def _ref_window_attention(qkv, table, B, R, Cdim, nH, shift):
    """Oracle attention core on token-order qkv [B*T, 3C]: roll -> partition -> attention -> reverse -> roll (htsat.py:452-474, :326-352)."""
    hd = Cdim // nH
    T = R * R
    x = qkv.view(B, R, R, 3 * Cdim)
    sh = shift if R > 8 else 0
    if sh > 0:
        x = torch.roll(x, shifts=(-sh, -sh), dims=(1, 2))
        mask = O.shift_attn_mask(R, R, 8, sh)
    else:
        mask = None
    xw = O.window_partition(x, 8).view(-1, 64, 3 * Cdim)
    B_ = xw.shape[0]
    q, k, v = xw.reshape(B_, 64, 3, nH, hd).permute(2, 0, 3, 1, 4)
    attn = q @ k.transpose(-2, -1)
    idx = O.relative_position_index()
    attn = attn + table[idx.view(-1)].view(64, 64, -1).permute(2, 0, 1).unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = (attn.view(B_ // nW, nW, nH, 64, 64) + mask.unsqueeze(1).unsqueeze(0)).view(-1, nH, 64, 64)
    attn = torch.softmax(attn, dim=-1)
    o = (attn @ v).transpose(1, 2).reshape(B_, 64, Cdim)
    o = O.window_reverse(o.view(-1, 8, 8, Cdim), 8, R, R)
    if sh > 0:
        o = torch.roll(o, shifts=(sh, sh), dims=(1, 2))
    return o.reshape(B * T, Cdim), attn


def _rand_qkv(B, R, Cdim, nH, seed):
    g = torch.Generator().manual_seed(seed)
    hd = Cdim // nH
    qkv = torch.randn(B * R * R, 3 * Cdim, generator=g)
    qkv[:, :Cdim] *= hd ** -0.5           # q pre-scaled (the C path folds the scale into the qkv weights)
    return bf16r(qkv), 0.5 * torch.randn(225, nH, generator=g), g


def check_window_attention(B, R, Cdim, nH, shift, seed=0):
    """ard_window_attention vs the oracle's attention core on identical (bf16-rounded) qkv."""
    lib = L.load()
    T = R * R
    qkv, table, _ = _rand_qkv(B, R, Cdim, nH, seed)
    ref_out, attn = _ref_window_attention(qkv, table, B, R, Cdim, nH, shift)
    qd = qkv.to("cuda", torch.bfloat16).contiguous()
    out = torch.zeros(B * T, Cdim, device="cuda", dtype=torch.bfloat16)
    cap = torch.zeros(attn.shape[0], nH, 64, 64, device="cuda", dtype=torch.float32)
    td = table.cuda().contiguous()
    L.check(lib.ard_window_attention(L.ptr(qd), L.ptr(out), L.ptr(td), L.ptr(cap), 1.0, 0, B, R, R, Cdim, nH, shift, L.stream_ptr()))
    torch.cuda.synchronize()
    return rel(out.float().cpu(), ref_out), rel(cap.cpu(), attn)


def check_window_attention_bwd(B, R, Cdim, nH, shift, seed=0):
    """ard_window_attention_bwd vs autograd through the oracle's attention core (same bf16-rounded qkv and dout)."""
    lib = L.load()
    T = R * R
    qkv, table, g = _rand_qkv(B, R, Cdim, nH, seed)
    dout = bf16r(torch.randn(B * T, Cdim, generator=g))
    qr = qkv.clone().requires_grad_(True)
    out, _ = _ref_window_attention(qr, table, B, R, Cdim, nH, shift)
    out.backward(dout)
    qd = qkv.to("cuda", torch.bfloat16).contiguous()
    dd = dout.to("cuda", torch.bfloat16).contiguous()
    dq = torch.zeros(B * T, 3 * Cdim, device="cuda", dtype=torch.bfloat16)
    td = table.cuda().contiguous()
    L.check(lib.ard_window_attention_bwd(L.ptr(qd), L.ptr(dd), L.ptr(dq), L.ptr(td), B, R, R, Cdim, nH, shift, L.stream_ptr()))
    torch.cuda.synchronize()
    got, ref = dq.float().cpu(), qr.grad
    return tuple(rel(got[:, i * Cdim:(i + 1) * Cdim], ref[:, i * Cdim:(i + 1) * Cdim]) for i in range(3))


def check_layernorm_bwd(rows, Cdim, seed=0, with_add=True):
    """ard_layernorm_bwd vs autograd of F.layer_norm."""
    lib = L.load()
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(rows, Cdim, generator=g) * 2 + 0.5).requires_grad_(True)
    gamma, beta = 1 + 0.2 * torch.randn(Cdim, generator=g), 0.1 * torch.randn(Cdim, generator=g)
    go, add = torch.randn(rows, Cdim, generator=g), torch.randn(rows, Cdim, generator=g)
    torch.nn.functional.layer_norm(x, (Cdim,), gamma, beta, 1e-5).backward(go)
    ref = x.grad + (add if with_add else 0)
    xd, gd, gm, ad = x.detach().cuda(), go.cuda(), gamma.cuda(), add.cuda()
    out = torch.empty(rows, Cdim, device="cuda")
    L.check(lib.ard_layernorm_bwd(L.ptr(xd), L.ptr(gd), L.ptr(gm), L.ptr(ad) if with_add else None, L.ptr(out), rows, Cdim, L.stream_ptr()))
    torch.cuda.synchronize()
    return rel(out.cpu(), ref)


def make_encoder(model="tiny", seed=0, fusion=False, residual=False):
    from audio_residual_b200.clap import build_clap_module
    from audio_residual_b200.residual import inject_residuals
    sd = W.make_state_dict(model, seed=seed)
    clap = build_clap_module(model, sd, device="cuda:0", enable_fusion=fusion)
    ores = None
    if residual:
        pca, lam = W.make_pca(model, seed=seed)
        if residual is not True:   # an iterable of layer indices
            pca = {l: pca[l] for l in residual}
        clap._residuals = inject_residuals(clap.model.audio_branch, pca, lam)
        ores = {l: (torch.tensor(pca[l]["mean"], dtype=torch.float32), torch.tensor(pca[l]["components"], dtype=torch.float32),
                    torch.from_numpy(lam[l])) for l in pca}
    return clap, sd, ores


def check_logmel(B=2):
    """ard_logmel (FFT + banded mel) vs the oracle's conv-DFT + dense mel matmul, before and after bn0."""
    clap, sd, _ = make_encoder("tiny")
    enc = clap.model.audio_branch
    h = enc._handle()
    wave = W.make_clips(B, seed=1234)
    out = torch.empty(B, 1001, 64, device="cuda")
    out_bn = torch.empty_like(out)
    lib = L.load()
    wd = wave.cuda()
    L.check(lib.ard_logmel(h, L.ptr(wd), B, 480000, 0, 0, L.ptr(out), L.stream_ptr()))
    L.check(lib.ard_logmel(h, L.ptr(wd), B, 480000, 1, 0, L.ptr(out_bn), L.stream_ptr()))
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = O.logmel(O.stft_power(wave, sd), sd)[:, 0]
        ref_bn = O.bn0_eval(ref, sd)
    return {"logmel": rel(out.cpu(), ref), "logmel_maxabs_dB": (out.cpu() - ref).abs().max().item(), "logmel_bn": rel(out_bn.cpu(), ref_bn)}


def check_fusion_featuriser(B=2):
    """ard_fusion_mel (device get_mel, data.py:363-399) vs the golden mel_fusion the reference's torchaudio featuriser produced,
    and the base+fusion encoder driven from the WAVEFORM through the public API vs the golden audio embedding."""
    g = np.load(os.path.join(GOLDEN, "htsat_base_fusion_b2.npz"))
    clap, sd, _ = make_encoder("base", seed=int(g["meta_seed"]), fusion=True)
    wave = W.make_clips(B, seed=1234)
    mf = clap.fusion_mel(wave.cuda())
    torch.cuda.synchronize()
    m = {"mel_fusion": rel(golden_sample(mf), torch.from_numpy(g["mel_fusion_sample"])),
         "channels_equal": float((mf[:, 0] - mf[:, 3]).abs().max().item())}
    emb = clap.get_audio_embedding_from_data(wave, use_tensor=True)
    m["audio_embed_from_waveform"] = rel(emb, torch.from_numpy(g["plain_audio_embed"]))
    return m


def encoder_outputs(clap, wave=None, mel_fusion=None):
    enc = clap.model.audio_branch
    if mel_fusion is not None:
        out = enc.encode(mel_fusion=mel_fusion.cuda(), want_dict=True, want_audio_embed=True)
    else:
        out = enc.encode(waveform=wave.cuda(), want_dict=True, want_audio_embed=True)
    torch.cuda.synchronize()
    return out


def compare_output_dicts(got, ref, ref_emb=None):
    m = {}
    for k in ("embedding", "clipwise_output", "framewise_output", "fine_grained_embedding"):
        m[k] = rel(got[k], ref[k])
    for l in range(4):
        m[f"res{l}"] = rel(got["layers_residuals"][l], ref["layers_residuals"][l])
        m[f"attn{l}"] = rel(got["layers_attention"][l], ref["layers_attention"][l])
    if ref_emb is not None:
        m["audio_embed"] = rel(got["audio_embed"], ref_emb)
    return m


def check_encoder_vs_oracle(model="tiny", B=2, residual=False, seed=0):
    clap, sd, ores = make_encoder(model, seed=seed, residual=residual)
    wave = W.make_clips(B, seed=1234)
    got = encoder_outputs(clap, wave)
    with torch.no_grad():
        ref = O.htsat_forward({"waveform": wave}, sd, O.CONFIGS[model], ores)
        ref_emb = O.audio_projection(ref["embedding"], sd)
    return compare_output_dicts(got, ref, ref_emb)


def check_encoder_vs_golden(fname="htsat_tiny_b2.npz"):
    g = np.load(os.path.join(GOLDEN, fname))
    model, seed, B, fusion = str(g["meta_model"]), int(g["meta_seed"]), int(g["meta_B"]), bool(int(g["meta_fusion"]))
    m = {}
    for tag, residual in (("plain", False), ("residual", True)):
        clap, sd, ores = make_encoder(model, seed=seed, fusion=fusion, residual=residual)
        wave = W.make_clips(B, seed=1234)
        if fusion:
            # the golden mel_fusion came from the reference's torchaudio featuriser; regenerate it with the oracle port
            fb = torch.from_numpy(W.mel_filterbank(htk=True, slaney_norm=False)).float()
            win = torch.from_numpy(W.hann_periodic(1024)).float()
            mel = torch.stack([O.fusion_mel(w, fb, win) for w in wave])
            m[f"{tag}_mel_fusion_input"] = rel(golden_sample(torch.stack([mel] * 4, dim=1)), torch.from_numpy(g["mel_fusion_sample"]))
            got = encoder_outputs(clap, mel_fusion=torch.stack([mel] * 4, dim=1))
        else:
            got = encoder_outputs(clap, wave)
        m[f"{tag}_embedding"] = rel(got["embedding"], torch.from_numpy(g[f"{tag}_embedding"]))
        m[f"{tag}_audio_embed"] = rel(got["audio_embed"], torch.from_numpy(g[f"{tag}_audio_embed"]))
        m[f"{tag}_clipwise"] = rel(got["clipwise_output"], torch.from_numpy(g[f"{tag}_clipwise_output"]))
        m[f"{tag}_framewise"] = rel(golden_sample(got["framewise_output"]), torch.from_numpy(g[f"{tag}_framewise_sample"]))
        m[f"{tag}_fine"] = rel(golden_sample(got["fine_grained_embedding"]), torch.from_numpy(g[f"{tag}_fine_sample"]))
        for l in range(4):
            m[f"{tag}_res{l}"] = rel(golden_sample(got["layers_residuals"][l]), torch.from_numpy(g[f"{tag}_res{l}_sample"]))
            m[f"{tag}_attn{l}"] = rel(golden_sample(got["layers_attention"][l]), torch.from_numpy(g[f"{tag}_attn{l}_sample"]))
    return m


def _train_step(clap, wave, text, labels, zero=True):
    """src/training.py:21-32 on the mirror API: returns (loss, sims); lambda gradients land in res.learnable.grad."""
    for r in clap._residuals.values():
        if zero:
            r.learnable.grad = None
    emb = clap.get_audio_embedding_from_data(wave.cuda(), use_tensor=True).float()
    sims = emb @ text.T.cuda()
    loss = torch.nn.CrossEntropyLoss()(sims, labels.cuda())
    loss.backward()
    torch.cuda.synchronize()
    return loss.item(), sims.detach().cpu()


def check_training_step_vs_golden(fname="htsat_tiny_b2.npz"):
    """One zero-shot training step (config c3): loss, similarities and d loss / d lambda of every layer vs what the
    reference's loss.backward() produced (tests/golden, generated by oracle/make_golden.py)."""
    g = np.load(os.path.join(GOLDEN, fname))
    model, seed, B = str(g["meta_model"]), int(g["meta_seed"]), int(g["meta_B"])
    clap, sd, _ = make_encoder(model, seed=seed, residual=True)
    wave = W.make_clips(B, seed=1234)
    text = W.make_text_embeds(50, 512, seed=7)
    labels = torch.from_numpy(g["labels"])
    loss, sims = _train_step(clap, wave, text, labels)
    m = {"loss_abs": abs(loss - float(g["train_loss"])), "sims": rel(sims, torch.from_numpy(g["train_sims"]))}
    for l, r in clap._residuals.items():
        m[f"lambda_grad{l}"] = rel(r.learnable.grad.cpu(), torch.from_numpy(g[f"lambda_grad{l}"]))
    return m


def check_training_step_vs_oracle(model="tiny", B=2, layers=(1,), seed=0, cosine=False):
    """ResiDual on a subset of layers (plain blocks above the patched ones exercise the un-patched block backward):
    lambda gradients vs autograd through the oracle."""
    clap, sd, ores = make_encoder(model, seed=seed, residual=tuple(layers))
    wave = W.make_clips(B, seed=99)
    text = W.make_text_embeds(50, 512, seed=7)
    labels = torch.from_numpy(np.random.default_rng(5).integers(0, 50, size=B))
    loss, sims = _train_step(clap, wave, text, labels)
    ores = {l: (mu, comp, lam.clone().requires_grad_(True)) for l, (mu, comp, lam) in ores.items()}
    oloss, osims = O.zero_shot_loss(wave, labels, text, sd, W.CONFIGS[model], ores)
    oloss.backward()
    m = {"loss_abs": abs(loss - oloss.item()), "sims": rel(sims, osims.detach())}
    for l, r in clap._residuals.items():
        a, b = r.learnable.grad.cpu().double(), ores[l][2].grad.double()
        m[f"lambda_grad{l}"] = rel(a, b)
        if cosine:
            m[f"lambda_cos{l}"] = (a @ b / (a.norm() * b.norm())).item()
            m[f"lambda_norm_ratio{l}"] = (a.norm() / b.norm()).item()
    return m


def check_embedding_grad_vs_oracle(model="tiny", B=2, layers=(0, 1, 2, 3), seed=0, wseed=99):
    """d loss / d lambda for a loss on output_dict['embedding'] (a linear probe on the 768-d embedding, no ReLU between the
    encoder and the loss, so the comparison is not perturbed by ReLU gates flipping under bf16 forward error)."""
    clap, sd, ores = make_encoder(model, seed=seed, residual=tuple(layers))
    enc = clap.model.audio_branch
    wave = W.make_clips(B, seed=wseed)
    NF = enc.num_features
    g = torch.Generator().manual_seed(3)
    Wc = torch.randn(50, NF, generator=g) / NF ** 0.5
    labels = torch.from_numpy(np.random.default_rng(5).integers(0, 50, size=B))
    for r in clap._residuals.values():
        r.learnable.grad = None
    emb = enc.encode(waveform=wave.cuda())["embedding"]
    loss = torch.nn.functional.cross_entropy(emb @ Wc.T.cuda(), labels.cuda())
    loss.backward()
    torch.cuda.synchronize()
    ores = {l: (mu, comp, lam.clone().requires_grad_(True)) for l, (mu, comp, lam) in ores.items()}
    oemb = O.htsat_forward({"waveform": wave}, sd, W.CONFIGS[model], ores)["embedding"]
    oloss = torch.nn.functional.cross_entropy(oemb @ Wc.T, labels)
    oloss.backward()
    m = {"loss_abs": abs(loss.item() - oloss.item()), "embedding": rel(emb.detach().cpu(), oemb.detach())}
    for l, r in clap._residuals.items():
        m[f"lambda_grad{l}"] = rel(r.learnable.grad.cpu(), ores[l][2].grad)
    return m


def check_stats(rows, D, strided=False, seed=0, calls=1):
    """ard_stats_accumulate(_strided) vs float64 X^T X / column sums (fp32-grade products are expected: split-bf16 GEMM)."""
    from audio_residual_b200.residual import MomentAccumulator
    g = torch.Generator().manual_seed(seed)
    nh = 3 if strided else 1
    full = torch.randn(rows, nh, D, generator=g) * 0.7 + 0.3
    xd = full.cuda()
    acc = MomentAccumulator(D, xd.device)
    view = xd[:, 1] if strided else xd[:, 0]
    per = (rows + calls - 1) // calls
    for c in range(calls):
        acc.update(view[c * per:(c + 1) * per])
    torch.cuda.synchronize()
    x64 = (full[:, 1] if strided else full[:, 0]).double()
    return {"n": acc.n - rows, "sum": rel(acc.s1.cpu(), x64.sum(0)), "sumsq": rel(acc.s2.cpu(), x64.t() @ x64)}


def check_pca_moments_vs_oracle(layer=2, B=2):
    """compute_pca_components' statistics of one layer's residuals (a19): GPU forward + tensor-core moments vs the oracle's
    residuals in float64; the eigen-spectrum of the two covariances is compared as well."""
    from audio_residual_b200.residual import MomentAccumulator
    clap, sd, _ = make_encoder("tiny")
    enc = clap.model.audio_branch
    wave = W.make_clips(B, seed=77)
    out = enc.encode(waveform=wave.cuda(), quantize=True, want_dict=True)
    res = out["layers_residuals"][layer]
    acc = MomentAccumulator(res.shape[-1], res.device)
    acc.update(res)
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = O.htsat_forward({"waveform": O.quantize_tensor(wave)}, sd, W.CONFIGS["tiny"])["layers_residuals"][layer]
    X = ref.reshape(-1, ref.shape[-1]).double()
    n = X.shape[0]
    mean_ref, mean = X.mean(0), acc.s1.cpu() / n
    cov_ref = (X - mean_ref).t() @ (X - mean_ref) / (n - 1)
    cov = (acc.s2.cpu() - n * torch.outer(mean, mean)) / (n - 1)
    ev_ref, ev = torch.linalg.eigvalsh(cov_ref).flip(0), torch.linalg.eigvalsh(cov).flip(0)
    k = min(32, ev.numel())
    return {"n": acc.n - n, "mean": rel(mean, mean_ref), "cov": rel(cov, cov_ref), "top_eigenvalues": rel(ev[:k], ev_ref[:k])}


def check_quantize_waveform(n=480000 * 2 + 8, seed=0):
    """ard_quantize_waveform vs the oracle's quantize_tensor (src/residual.py:210-212): integer arithmetic, bit-exact."""
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(n, generator=g) * 2.4 - 1.2)          # includes values outside [-1, 1]
    x[:6] = torch.tensor([1.0, -1.0, 0.0, 1.0 / 32767.0, -0.5 / 32767.0, 0.99999])
    xd = x.cuda()
    out = torch.empty_like(xd)
    L.check(L.load().ard_quantize_waveform(L.ptr(xd), L.ptr(out), n, L.stream_ptr()))
    torch.cuda.synchronize()
    return int((out.cpu() != O.quantize_tensor(x)).sum().item())


def check_argmax_vs_golden(fname="htsat_tiny_b2.npz"):
    """north_star: argmax class predictions identical to the reference's (zero-shot similarities over 50 classes with the
    ResiDual-patched model, and the 527-way clipwise output of the plain model)."""
    g = np.load(os.path.join(GOLDEN, fname))
    model, seed, B = str(g["meta_model"]), int(g["meta_seed"]), int(g["meta_B"])
    wave = W.make_clips(B, seed=1234)
    text = W.make_text_embeds(50, 512, seed=7)
    clap, _, _ = make_encoder(model, seed=seed, residual=True)
    with torch.no_grad():
        emb = clap.get_audio_embedding_from_data(wave.cuda(), use_tensor=True).float().cpu()
    sims = emb @ text.T
    ref_sims = torch.from_numpy(g["train_sims"])
    clap2, _, _ = make_encoder(model, seed=seed)
    with torch.no_grad():
        clip = encoder_outputs(clap2, wave)["clipwise_output"]
    ref_clip = torch.from_numpy(g["plain_clipwise_output"])
    margin = (ref_sims.topk(2, dim=-1).values[:, 0] - ref_sims.topk(2, dim=-1).values[:, 1]).min().item()
    return {"zero_shot_argmax_mismatch": int((sims.argmax(-1) != ref_sims.argmax(-1)).sum()),
            "clipwise_argmax_mismatch": int((clip.argmax(-1).cpu() != ref_clip.argmax(-1)).sum()),
            "reference_top1_top2_margin": margin, "sims_max_abs_err": (sims - ref_sims).abs().max().item()}


def check_attention_block(B=2, block=0, residual=False, seed=0):
    """ard_attention_block (LayerNorm1 + qkv + window attention + (folded) projection + shortcut in one tcgen05 kernel) vs the
    oracle's x + residual_x of the same block (htsat.py:449-476, src/residual.py:58-92), and vs the unfused kernel chain."""
    clap, sd, ores = make_encoder("tiny", seed=seed, residual=residual)
    enc = clap.model.audio_branch
    h = enc._handle()
    l, R, Cd = 0, 64, 96
    x = torch.randn(B, R * R, Cd, generator=torch.Generator().manual_seed(100 + block)) * 0.8 + 0.1
    xd = x.cuda().contiguous()
    out = torch.empty_like(xd)
    L.check(L.load().ard_attention_block(h, l, block, L.ptr(xd), B, L.ptr(out), L.stream_ptr()))
    torch.cuda.synchronize()
    _, _, res_unfused = enc.layers[l].blocks[block](xd)                       # ard_block_forward: unfused chain (capture path)
    with torch.no_grad():
        _, _, o_res = O.swin_block(x, sd, f"layers.{l}.blocks.{block}.", R, R, enc.num_heads[l], 0 if block % 2 == 0 else 4,
                                   ores[l] if residual else None)
    got_r = out.cpu() - x
    return {"out": rel(out, x + o_res), "branch": rel(got_r, o_res), "branch_vs_unfused": rel(got_r, res_unfused.cpu())}
