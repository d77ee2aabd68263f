"""Mirror of the CLAP wrapper API the Audio-ResiDual drivers call, on top of libard_b200.so.

* CLAP           <- clap_module/model.py (audio side only): encode_audio :589-590, audio_projection :539-543,
                    get_audio_embedding :720-742, get_audio_output_dict :745-762.
* CLAP_Module    <- hook.py: get_audio_embedding_from_data(x, use_tensor=False, data_fil="repeatpad") :158-192.
* get_audio_features / batch_features <- training/data.py:402-506 (the reachable <= max_len branches), batched on device
  instead of the reference's per-clip Python loop.
Text towers, checkpoint download and tokenisers are out of scope (they need the network and are not on the path).
"""
import functools
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import weights as W
from .htsat import HTSAT_Swin_Transformer, create_htsat_model

AUDIO_CFG = {"audio_length": 1024, "clip_samples": 480000, "mel_bins": 64, "sample_rate": 48000, "window_size": 1024,
             "hop_size": 480, "fmin": 50, "fmax": 14000, "class_num": 527, "model_type": "HTSAT"}


def int16_to_float32(x):
    """data.py:93-94"""
    return (x / 32767.0).astype("float32")


def float32_to_int16(x):
    """data.py:97-99"""
    x = np.clip(x, a_min=-1.0, a_max=1.0)
    return (x * 32767.0).astype("int16")


def _fill(wave, max_len, data_filling):
    """data.py:469-496 for one clip (1-D tensor) with len <= max_len."""
    n = wave.shape[0]
    if n == max_len:
        return wave
    if data_filling == "repeatpad":
        wave = wave.repeat(int(max_len / n))
        return F.pad(wave, (0, max_len - wave.shape[0]), mode="constant", value=0)
    if data_filling == "pad":
        return F.pad(wave, (0, max_len - n), mode="constant", value=0)
    if data_filling == "repeat":
        return wave.repeat(int(max_len / n) + 1)[:max_len]
    raise NotImplementedError(f"data_filling {data_filling} not implemented")


def get_audio_features(sample, audio_data, max_len, data_truncating, data_filling, audio_cfg, require_grad=False):
    """data.py:402-506 restricted to what the reference can reach: clips longer than max_len crash there
    (np.random.integers does not exist, data.py:467), so they raise here too. `mel_fusion` is produced by the encoder's
    on-device featuriser (CLAP_Module.fusion_mel), not per clip."""
    if data_truncating not in ("rand_trunc", "fusion"):
        raise NotImplementedError(f"data_truncating {data_truncating} not implemented")
    if len(audio_data) > max_len:
        raise AttributeError("module 'numpy.random' has no attribute 'integers' (reference data.py:467): clips longer than "
                             "max_len are unreachable")
    sample["waveform"] = _fill(audio_data, max_len, data_filling)
    sample["longer"] = torch.tensor([False])
    return sample


_FILL_MODES = {"repeatpad": 0, "pad": 1, "repeat": 2}


def _fill_on_device(flat, offsets, lengths, B, max_len, mode, quantize=False):
    """ard_fill_clips: `flat` (CUDA, fp32 or int16 PCM) holds the clips back to back -> [B, max_len] fp32 on the same device."""
    import ctypes as C
    from . import lib as L
    out = torch.empty((B, max_len), device=flat.device, dtype=torch.float32)
    with torch.cuda.device(flat.device):
        L.check(L.load().ard_fill_clips(C.c_void_p(flat.data_ptr()), int(flat.dtype == torch.int16), L.ptr(offsets), L.ptr(lengths), B, max_len, mode,
                                        int(bool(quantize)), L.ptr(out), L.stream_ptr()))
    return out


def batch_features(x, max_len=480000, data_filling="repeatpad", device=None, do_pad_or_truncate=False, quantize=False):
    """Batched get_audio_features (data.py:402-506, the reachable len <= max_len branches): x is [B, T] (tensor / ndarray) or
    a list of 1-D clips of different lengths, float32 or int16 PCM (int16 is read as int16_to_float32, data.py:93-94).
    Returns a [B, max_len] float32 tensor on `device`. With a CUDA `device` the clips travel as ONE flat buffer and the
    repeat / pad filling (and the optional int16 round trip, `quantize`) runs in one kernel (ard_fill_clips) instead of the
    reference's per-clip Python loop (hook.py:175-188); without a device (host-only callers) the same rule is applied per clip."""
    if data_filling not in _FILL_MODES:
        raise NotImplementedError(f"data_filling {data_filling} not implemented")           # data.py:492-496
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    if torch.is_tensor(x) and x.dim() == 2:
        clips = None if x.shape[1] == max_len and not do_pad_or_truncate else [c for c in x]
    else:
        clips = [torch.as_tensor(c) for c in x]
    dev = torch.device(device) if device is not None else None
    if clips is None:                                    # already [B, max_len]
        if x.dtype == torch.int16 or quantize:
            if dev is None or dev.type != "cuda":
                y = x.float() / 32767.0 if x.dtype == torch.int16 else x.float()
                if quantize:
                    from .residual import quantize_tensor
                    y = quantize_tensor(y)
                return y if dev is None else y.to(dev)
            src = x.to(dev) if x.dtype == torch.int16 else x.to(dev, torch.float32)
            return _fill_on_device(src.contiguous(), None, None, x.shape[0], max_len, 0, quantize)
        return x.to(device=dev, dtype=torch.float32) if dev is not None else x.float()
    if do_pad_or_truncate:
        from .residual import pad_or_truncate
        clips = [pad_or_truncate(c if c.dtype != torch.int16 else c.float() / 32767.0, max_len) for c in clips]
    for c in clips:
        if c.dim() != 1:
            raise ValueError(f"each clip must be 1-D (got shape {tuple(c.shape)})")
        if c.shape[0] > max_len:
            raise AttributeError("clips longer than max_len are unreachable in the reference (data.py:467)")
        if c.shape[0] == 0:
            raise ZeroDivisionError("division by zero")      # int(max_len / len(audio_data)), data.py:472
    if dev is None or dev.type != "cuda":
        out = torch.stack([_fill(c.float() / 32767.0 if c.dtype == torch.int16 else c.float(), max_len, data_filling) for c in clips])
        if quantize:
            from .residual import quantize_tensor
            out = quantize_tensor(out)
        return out if dev is None else out.to(dev)
    pcm = all(c.dtype == torch.int16 for c in clips)
    clips = [c if pcm else (c.float() / 32767.0 if c.dtype == torch.int16 else c.to(torch.float32)) for c in clips]
    lengths = torch.tensor([c.shape[0] for c in clips], dtype=torch.int32)
    offsets = torch.zeros(len(clips), dtype=torch.int64)
    offsets[1:] = torch.cumsum(lengths[:-1].to(torch.int64), 0)
    flat = torch.cat([c.to(dev, non_blocking=True) for c in clips]) if any(c.is_cuda for c in clips) else torch.cat(clips).to(dev)
    return _fill_on_device(flat.contiguous(), offsets.to(dev), lengths.to(dev), len(clips), max_len, _FILL_MODES[data_filling], quantize)


class CLAP(nn.Module):
    """Audio side of clap_module/model.py::CLAP."""

    def __init__(self, embed_dim, audio_cfg, text_cfg=None, enable_fusion=False, fusion_type="None", joint_embed_shape=512,
                 mlp_act="relu"):
        super().__init__()
        self.audio_cfg, self.enable_fusion, self.fusion_type = audio_cfg, enable_fusion, fusion_type
        self.joint_embed_shape = joint_embed_shape
        if audio_cfg.get("model_type", "HTSAT") != "HTSAT":
            raise RuntimeError(f"Model config for {audio_cfg.get('model_type')} not found.")   # model.py:469-470
        if mlp_act != "relu":
            raise NotImplementedError("audio_projection activation other than ReLU")
        self.audio_branch = create_htsat_model(audio_cfg, enable_fusion, fusion_type)
        self.audio_projection = nn.Sequential(nn.Linear(embed_dim, joint_embed_shape), nn.ReLU(),
                                              nn.Linear(joint_embed_shape, joint_embed_shape))
        for p in self.parameters():
            p.requires_grad = False

    def __setattr__(self, name, value):
        super().__setattr__(name, value)
        if name in ("audio_branch", "audio_projection"):    # callers re-assign audio_branch (src/training.py:103)
            ab = self._modules.get("audio_branch")
            pj = self._modules.get("audio_projection")
            if ab is not None and pj is not None:
                object.__setattr__(ab, "_projection", pj)

    def _input(self, data):
        keys = data[0].keys()
        return {k: torch.cat([d[k].unsqueeze(0) for d in data], dim=0) for k in keys}   # model.py:735-738

    def encode_audio(self, audio, device=None):
        return self.audio_branch(audio, mixup_lambda=None, device=device)

    def get_audio_embedding(self, data):
        """model.py:720-742: list of per-clip dicts (or an already batched dict) -> L2-normalised [N, joint] embeddings."""
        inp = data if isinstance(data, dict) else self._input(data)
        ab = self.audio_branch
        if ab.enable_fusion:
            out = ab.encode(mel_fusion=inp["mel_fusion"], want_audio_embed=True)
        else:
            out = ab.encode(waveform=inp["waveform"], want_audio_embed=True)
        return out["audio_embed"]

    def get_audio_output_dict(self, data):
        """model.py:745-762 (fork addition)."""
        inp = data if isinstance(data, dict) else self._input(data)
        return self.encode_audio(inp)


class CLAP_Module(nn.Module):
    """hook.py::CLAP_Module, audio entry points only."""

    def __init__(self, enable_fusion=False, device=None, amodel="HTSAT-tiny", tmodel="roberta"):
        super().__init__()
        if device is None:
            device = "cuda:0"
        name = amodel.split("-")[-1]
        if name not in W.CONFIGS:
            raise RuntimeError(f"Model config for {amodel} not found.")
        cfg = W.CONFIGS[name]
        audio_cfg = dict(AUDIO_CFG, model_name=name)
        self.enable_fusion = enable_fusion
        self.model_cfg = {"embed_dim": cfg["embed_dim"] * 8, "audio_cfg": audio_cfg}
        self.model = CLAP(cfg["embed_dim"] * 8, audio_cfg, enable_fusion=enable_fusion,
                          fusion_type="aff_2d" if enable_fusion else "None", joint_embed_shape=cfg["joint_dim"])
        self.device = torch.device(device)
        self.model.to(self.device)
        self._htk = None

    def load_state_dict_flat(self, sd):
        """Load a flat dict with un-prefixed audio_branch keys + audio_projection.* keys (weights.make_state_dict layout,
        i.e. the reference checkpoint's `audio_branch.` / `audio_projection.` tensors, factory.py:53-70)."""
        ab = self.model.audio_branch
        own = ab.state_dict()
        ab.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)
        self.model.audio_projection.load_state_dict({k[len("audio_projection."):]: v for k, v in sd.items()
                                                     if k.startswith("audio_projection.")})
        self.model.to(self.device)
        return self

    def load_ckpt(self, ckpt=None, model_id=-1, verbose=True):
        """hook.py:75-119 + factory.py:53-70 for the audio side. `ckpt` is a path to (or an already loaded) LAION-CLAP
        checkpoint: optional {"state_dict": ...} wrapper, optional "module." prefix, audio tower under `audio_branch.`, the
        projection under `audio_projection.`; text tower / logit scales / the unreachable aff_2d fusion branch are ignored.
        Downloading the paper checkpoints (ckpt=None) needs the network and is not available here."""
        if ckpt is None:
            raise RuntimeError("load_ckpt(ckpt=None) would download the LAION-CLAP weights (hook.py:91-112): pass a local checkpoint")
        sd = torch.load(ckpt, map_location="cpu", weights_only=False) if isinstance(ckpt, (str, bytes)) or hasattr(ckpt, "__fspath__") else ckpt
        if isinstance(sd, dict) and "state_dict" in sd:
            sd = sd["state_dict"]
        if next(iter(sd)).startswith("module"):
            sd = {k[7:]: v for k, v in sd.items()}
        ab = self.model.audio_branch
        own = ab.state_dict()
        branch = {k[len("audio_branch."):]: v for k, v in sd.items() if k.startswith("audio_branch.")}
        # index / mask / counter buffers are derived from the architecture: optional in the checkpoint
        derived = ("relative_position_index", "attn_mask", "num_batches_tracked")
        missing = [k for k in own if k not in branch and not k.endswith(derived)]
        if missing:
            raise RuntimeError(f"Error(s) in loading state_dict for {type(ab).__name__}: Missing key(s): {missing[:8]}"
                               f"{' ...' if len(missing) > 8 else ''}")
        ab.load_state_dict({k: branch[k] for k in own if k in branch}, strict=False)
        proj = {k[len("audio_projection."):]: v for k, v in sd.items() if k.startswith("audio_projection.")}
        self.model.audio_projection.load_state_dict(proj)
        self.model.to(self.device)
        if verbose:
            skipped = sorted({k.split(".")[0] for k in sd if not k.startswith(("audio_branch.", "audio_projection."))})
            print(f"Loaded {len(own)} audio_branch and {len(proj)} audio_projection tensors; ignored groups: {skipped}")
        return self

    def fusion_mel(self, wave, quantize=False):
        """get_mel (data.py:363-399) for a batch on device, stacked 4x as data.py:497-501 does for clips <= 10 s."""
        return self.model.audio_branch.fusion_mel(wave, quantize=quantize)

    def get_audio_embedding_from_data(self, x, use_tensor=False, data_fil="repeatpad"):
        """hook.py:158-192. use_tensor=False: numpy/tensor input, int16 round-trip first, returns numpy.
        use_tensor=True: tensor input, no quantisation, returns a tensor.
        Extension: int16 input (numpy / tensor, PCM samples as a wav file stores them) is read as int16_to_float32(x)
        (data.py:93-94) on the device, so it crosses PCIe at 0.96 MB per clip instead of 1.92 MB; results are bit-identical
        to passing that float array."""
        self.model.eval()
        enc = self.model.audio_branch
        training_step = torch.is_grad_enabled() and any(p is not None and p.requires_grad for p in enc._lambda_params())
        if isinstance(x, np.ndarray) and x.ndim == 2 and x.dtype in (np.float32, np.int16):
            xt = torch.from_numpy(x)            # zero-copy view: the numpy route rides the same chunked pipeline
        else:
            xt = x
        if (torch.is_tensor(xt) and xt.device.type == "cpu" and xt.dim() == 2 and xt.shape[1] == 480000
                and xt.dtype in (torch.float32, torch.int16) and xt.shape[0] > self.h2d_chunk and not training_step):
            emb = self._embed_host_pipelined(xt, quantize=not use_tensor)
        else:
            wave = batch_features(x, 480000, data_fil, device=self.device)
            if enc.enable_fusion:
                out = enc.encode(mel_fusion=self.fusion_mel(wave, quantize=not use_tensor), want_audio_embed=True)
            else:
                out = enc.encode(waveform=wave, quantize=not use_tensor, want_audio_embed=True)
            emb = out["audio_embed"]
        if not use_tensor:
            emb = emb.detach().cpu().numpy()
        return emb

    # ARD_PIPE_ADAPT: 0 = fixed schedule, no event timing; 1 (default) = time every call's copies / encodes and fit the pipeline model
    # (bench.py reports the fit per rank), keep the fixed schedule; 2 = also re-plan the chunk sizes from the fit. Re-planning is
    # opt-in: on the 8-GPU box it was measured on, the slow ranks were not copy-bound (their copies ran at 35-43 GB/s inside the
    # pipeline) and the re-planned ranks were no faster (15.4 vs 15.0 ms per call), while each re-plan costs an eager + a capture call.
    h2d_adapt = int(os.environ.get("ARD_PIPE_ADAPT", "1") or 0)
    h2d_chunk = 64                        # host batches larger than this are copied in chunks overlapped with the encoder
    h2d_schedule = (24, 50, 80, 116, 156, 204, 256)   # chunk sizes in clips: a small first copy (nothing overlaps it), then growing
    h2d_schedule_pcm16 = (32, 80, 144, 256)            # int16 transport: copies take half as long, so chunks may grow faster

    def _chunk_bounds(self, N, schedule=None):
        """Each copy must fit under the previous chunk's encode: 0.0346 ms/clip over PCIe gen5 (55.5 GB/s measured) against
        0.79 ms + 0.040 ms/clip of graph-replayed encoder time on a B200 (tools/batch_sweep.py), i.e. the next chunk may hold
        at most 22.8 + 1.156 x the clips of the current one; chunks are capped so the two staging buffers stay small for any N."""
        schedule = schedule or self.h2d_schedule
        bounds, lo, i = [], 0, 0
        while lo < N:
            c = schedule[min(i, len(schedule) - 1)]
            hi = min(N, lo + c)
            if N - hi < c // 3:      # do not leave a tiny tail chunk
                hi = N
            bounds.append((lo, hi))
            lo, i = hi, i + 1
        return bounds

    # ---- adaptive schedule. The fixed schedules above assume this GPU has the host to itself (55 GB/s). With 8 ranks on one socket
    # the pinned-host copy rate per GPU was measured at 23-36 GB/s: copies no longer fit under the previous chunk's encode, the
    # encoder stalls between chunks, and the call ends one encode of a LARGE last chunk after the last byte arrives (8 GPUs:
    # e2e 14.8 ms per 256 clips against 11.9 ms alone). So every pipelined call times its own copies and encodes with CUDA events
    # (waits excluded), the next call fits  copy(n) = c n  and  encode(n) = a + b n  to them, simulates the two-stream pipeline
    # for a family of candidate schedules (first chunk, growth factor, optional tapered tail) and takes the fastest.
    @staticmethod
    def _simulate(sizes, c, a, b, pcm):
        """End time (ms) of the last encode. Copy k needs staging buffer k % 2: free once chunk k-2 has been expanded (int16: at
        the start of its encode) or encoded (fp32)."""
        copy_done, enc_start, enc_done = [], [], []
        for k, n in enumerate(sizes):
            t = copy_done[k - 1] if k else 0.0
            if k >= 2:
                t = max(t, (enc_start[k - 2] + 0.02) if pcm else enc_done[k - 2])
            copy_done.append(t + 0.01 + c * n)
            st = max(copy_done[k], enc_done[k - 1] if k else 0.0)
            enc_start.append(st)
            enc_done.append(st + a + b * n)
        return enc_done[-1]

    @staticmethod
    @functools.lru_cache(maxsize=32)
    def _candidates(N):
        out = []
        for first in (16, 24, 32, 48):
            for growth in (1.0, 1.4, 1.8, 2.4):
                for tail in ((), (40, 24), (64, 32), (24,)):
                    body = N - sum(tail)
                    if body < first:
                        continue
                    sizes, cur, left = [], float(first), body
                    while left > 0:
                        n = min(left, int(round(cur)))
                        if left - n < 12:          # no tiny remainder
                            n = left
                        sizes.append(n)
                        left -= n
                        cur = min(cur * growth if growth > 1.0 else 64.0, 160.0)
                    out.append(tuple(sizes) + tuple(tail))
        return tuple(out)

    def _pick_bounds(self, N, dtype):
        """The schedule planned for (N, dtype) by an earlier call's _plan_next, else the fixed compute-bound schedule."""
        plan = getattr(self, "_pipe_plan", {}).get((N, dtype))
        if plan is not None:
            return plan
        return self._chunk_bounds(N, self.h2d_schedule_pcm16 if dtype == torch.int16 else None)

    def _plan_next(self, N, dtype):
        """Fit the rates to the newest COMPLETED call's events and plan the next call's schedule. Called at the end of a pipelined
        call, when all of its work is queued: the planning (a few hundred Python steps) runs while the GPU is busy, not in front
        of the first copy."""
        pcm = dtype == torch.int16
        rates = getattr(self, "_pipe_rates", {}).get(dtype)
        if rates is None:
            return
        done = rates.get("done")
        if done is None or not all(e.query() for rec in done for e in rec[:2]):
            return
        rates["done"] = None
        # The first two calls on a schedule are not representative: a new chunk size runs kernel by kernel once and is captured
        # into a CUDA graph on its second use (ard_api.cu::forward_graphed). Fitting those would re-plan on garbage, and every
        # re-plan brings new chunk sizes: the schedule would never settle (measured: ranks that flapped spent 20 ms per call).
        key = (N, dtype)
        seen = rates.setdefault("calls_on_plan", {})
        seen[key] = seen.get(key, 0) + 1
        if seen[key] <= 2 or rates.setdefault("replans", {}).get(key, 0) >= 3:
            return
        cp = [(n, e0.elapsed_time(e1)) for e0, e1, kind, n in done if kind == "copy"]
        en = [(n, e0.elapsed_time(e1)) for e0, e1, kind, n in done if kind == "enc"]
        if not cp or not en:
            return
        rates["c"] = sum(t for _, t in cp) / sum(n for n, _ in cp)
        xs, ys = [float(n) for n, _ in en], [t for _, t in en]
        mx, my = sum(xs) / len(xs), sum(ys) / len(ys)
        sxx = sum((x - mx) ** 2 for x in xs)
        bb = sum((x - mx) * (y - my) for x, y in zip(xs, ys)) / sxx if sxx > 0 else 0.0
        if bb <= 0.0 or my - bb * mx < 0.0:          # one chunk size, or noise: keep the slope of a proportional model
            bb, aa = 0.9 * my / mx, 0.1 * my
        else:
            aa = my - bb * mx
        rates["a"], rates["b"] = aa, bb
        c, a, b = rates["c"], aa, bb
        default = self._chunk_bounds(N, self.h2d_schedule_pcm16 if pcm else None)
        cur = [hi - lo for lo, hi in self._pick_bounds(N, dtype)]
        cur_t = self._simulate(cur, c, a, b, pcm)
        best, best_t = cur, cur_t
        for sizes in ([hi - lo for lo, hi in default],) + self._candidates(N):
            t = self._simulate(list(sizes), c, a, b, pcm)
            if t < best_t:
                best, best_t = list(sizes), t
        rates["predicted_ms"] = cur_t
        if self.h2d_adapt < 2 or best_t >= 0.95 * cur_t:                             # re-plan only for a real gain: new chunk sizes cost an eager run + a capture
            return
        rates["predicted_ms"] = best_t
        bounds, lo = [], 0
        for n in best:
            bounds.append((lo, lo + n))
            lo += n
        if not hasattr(self, "_pipe_plan"):
            self._pipe_plan = {}
        self._pipe_plan[key] = bounds
        seen[key] = 0
        rates["replans"][key] = rates["replans"].get(key, 0) + 1

    def _embed_host_pipelined(self, x, quantize):
        """Full-length host batch [N, 480000] fp32 or int16 PCM: copy chunk k+1 on a side stream while chunk k is encoded, so the
        PCIe transfer (1.92 / 0.96 MB per clip) hides behind compute instead of adding to it. Pinned input makes the copies
        asynchronous. Only the first copy is exposed, so the first chunk is small and later ones grow (_chunk_bounds).
        int16 chunks are expanded to fp32 on the device (ard_fill_clips) into a fixed buffer, which keeps the encoder's
        CUDA-graph key (input pointer, batch) stable across calls."""
        enc = self.model.audio_branch
        dev = self.device
        N = x.shape[0]
        pcm = x.dtype == torch.int16
        bounds = self._pick_bounds(N, x.dtype)
        self._last_bounds = bounds
        cmax = max(hi - lo for lo, hi in bounds)
        timing = []                       # (start, end, "copy" | "enc", clips) CUDA events of this call, fitted by a later call's _plan_next
        adapt = self.h2d_adapt            # ARD_PIPE_ADAPT=0: fixed schedule, no event timing (A/B measurements)
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream()
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(device=dev)
                self._stage = {}
            st = self._stage.get(x.dtype)
            if st is None or st[0].shape[0] < cmax:
                st = [torch.empty((cmax, 480000), device=dev, dtype=x.dtype) for _ in range(2)]
                self._stage[x.dtype] = st
                self._stage_free = {} if not hasattr(self, "_stage_free") else self._stage_free
                self._stage_free[x.dtype] = [torch.cuda.Event() for _ in range(2)]
            free = self._stage_free[x.dtype]
            if pcm and (getattr(self, "_pcm_wave", None) is None or self._pcm_wave.shape[0] < cmax):
                self._pcm_wave = torch.empty((cmax, 480000), device=dev, dtype=torch.float32)
            out = torch.empty((N, enc.joint_dim), device=dev, dtype=torch.float32)
            copied = [torch.cuda.Event() for _ in range(2)]
            import ctypes as C
            from . import lib as L
            lib = L.load()

            def start_copy(k):
                lo, hi = bounds[k]
                with torch.cuda.stream(self._copy_stream):
                    self._copy_stream.wait_event(free[k % 2])    # the encoder finished reading this staging buffer
                    if adapt:
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(self._copy_stream)
                    st[k % 2][:hi - lo].copy_(x[lo:hi], non_blocking=True)
                    if adapt:
                        e1.record(self._copy_stream)
                        timing.append((e0, e1, "copy", hi - lo))
                    copied[k % 2].record(self._copy_stream)

            for b in range(2):
                free[b].record(main)
            start_copy(0)
            if len(bounds) > 1:
                start_copy(1)
            # everything below this line runs while the first copy is in flight (the only exposed one): validating the handle
            # (weights / ResiDual signatures) costs ~0.3 ms of Python
            if not enc.enable_fusion:
                hnd = enc._handle()
                emb_scratch = torch.empty((cmax, enc.num_features), device=dev, dtype=torch.float32)
                fa = L.ArdForwardArgs()
                fa.quantize = int(bool(quantize))
                fa.precision = 1 if getattr(enc, "precision", "bf16") == "fp32" else 0
            for k, (lo, hi) in enumerate(bounds):
                if k >= 1 and k + 1 < len(bounds):
                    start_copy(k + 1)
                main.wait_event(copied[k % 2])
                if adapt:
                    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    t0.record(main)
                chunk = st[k % 2][:hi - lo]
                if pcm:   # int16_to_float32 on the device; the staging buffer is free again as soon as this kernel has run
                    wavef = self._pcm_wave[:hi - lo]
                    L.check(lib.ard_fill_clips(C.c_void_p(chunk.data_ptr()), 1, None, None, hi - lo, 480000, 0, 0, L.ptr(wavef), L.stream_ptr()))
                    free[k % 2].record(main)
                    chunk = wavef
                if enc.enable_fusion:   # get_mel + 4x stack on device (data.py:363-399, :497-501), then the fused encoder
                    res = enc.encode(mel_fusion=enc.fusion_mel(chunk, quantize=quantize), want_audio_embed=True)
                    out[lo:hi].copy_(res["audio_embed"])
                else:
                    # lean per-chunk call: the handle was validated once for this call (weights / ResiDuals cannot change inside
                    # it), the embedding lands straight in its rows of `out`. Keeps the host side of a chunk to two ctypes calls:
                    # on a loaded host (8 ranks on one socket) the Python around encode() was what the GPU waited for
                    fa.B, fa.waveform = hi - lo, chunk.data_ptr()
                    fa.embedding = emb_scratch.data_ptr()
                    fa.audio_embed = out.data_ptr() + lo * out.shape[1] * 4
                    L.check(lib.ard_encoder_forward(hnd, C.byref(fa), L.stream_ptr()))
                if adapt:
                    t1.record(main)
                    timing.append((t0, t1, "enc", hi - lo))
                if not pcm:
                    free[k % 2].record(main)
            if not hasattr(self, "_pipe_rates"):
                self._pipe_rates = {}
            if adapt:
                rates = self._pipe_rates.setdefault(x.dtype, {})
                self._plan_next(N, x.dtype)      # from the previous call's events (complete by now), while this call's work runs
                rates["done"] = timing
        return out


def build_clap_module(model="tiny", state_dict=None, device="cuda:0", enable_fusion=False, seed=0):
    """CLAP_Module with synthetic (or given) weights: the offline stand-in for hook.py's load_ckpt()."""
    m = CLAP_Module(enable_fusion=enable_fusion, device=device, amodel=f"HTSAT-{model}")
    sd = state_dict if state_dict is not None else W.make_state_dict(model, seed=seed)
    return m.load_state_dict_flat(sd)
