"""Thin ctypes binding of libdiffsal_b200's C ABI (include/diffsal_b200.h).

PyTorch is used only for device memory and streams; every FLOP of the denoiser runs in the
hand-written sm_100a kernels behind the handle.  No fallback: a missing library, a missing GPU or
an unsupported configuration raises.
"""
import ctypes

import torch

from . import _lib

OP_EVAL, OP_AXPY, OP_CLAMP, OP_DYNTHRESH, OP_POSTPROCESS = 0, 1, 2, 3, 4


class DsbConfig(ctypes.Structure):
    _fields_ = [("max_batch", ctypes.c_int), ("audio_visual", ctypes.c_int), ("reserved", ctypes.c_int * 6)]


class DsbSamplerOp(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("t", ctypes.c_float), ("dst", ctypes.c_int), ("nin", ctypes.c_int),
                ("src", ctypes.c_int * 4), ("coef", ctypes.c_float * 4), ("noise_coef", ctypes.c_float),
                ("noise_index", ctypes.c_int)]


class DsbSamplerDesc(ctypes.Structure):
    _fields_ = [("ops", ctypes.POINTER(DsbSamplerOp)), ("n_ops", ctypes.c_int), ("noise", ctypes.c_void_p),
                ("use_graph", ctypes.c_int), ("out_u8", ctypes.c_void_p)]


class DsbError(RuntimeError):
    pass


def _bind(lib):
    if getattr(lib, "_dsb_bound", False):
        return lib
    vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    lib.dsb_create.argtypes = [ctypes.POINTER(DsbConfig), ctypes.POINTER(vp)]
    lib.dsb_create.restype = ci
    lib.dsb_destroy.argtypes = [vp]
    lib.dsb_destroy.restype = None
    lib.dsb_last_error.argtypes = [vp]
    lib.dsb_last_error.restype = ctypes.c_char_p
    lib.dsb_workspace_bytes.argtypes = [vp, ci]
    lib.dsb_workspace_bytes.restype = ctypes.c_size_t
    lib.dsb_load_weight.argtypes = [vp, ctypes.c_char_p, vp, ctypes.POINTER(ctypes.c_int64), ci]
    lib.dsb_load_weight.restype = ci
    lib.dsb_finalize_weights.argtypes = [vp]
    lib.dsb_finalize_weights.restype = ci
    lib.dsb_set_condition.argtypes = [vp, ctypes.POINTER(vp), vp, ci, vp]
    lib.dsb_set_condition.restype = ci
    lib.dsb_condition_hash.argtypes = [vp, ctypes.POINTER(vp), vp, ci, ctypes.POINTER(ctypes.c_uint64), vp]
    lib.dsb_condition_hash.restype = ci
    lib.dsb_denoise.argtypes = [vp, vp, vp, vp, ci, vp]
    lib.dsb_denoise.restype = ci
    lib.dsb_sampler_update.argtypes = [vp, ctypes.POINTER(cf), ctypes.POINTER(vp), ci, vp, cf, vp, ctypes.c_int64, vp]
    lib.dsb_sampler_update.restype = ci
    lib.dsb_sample.argtypes = [vp, ctypes.POINTER(DsbSamplerDesc), vp, ci, vp]
    lib.dsb_sample.restype = ci
    lib.dsb_last_launch_count.argtypes = [vp]
    lib.dsb_last_launch_count.restype = ctypes.c_int64
    lib.dsb_debug_read.argtypes = [vp, ctypes.c_char_p, vp, ctypes.c_int64, vp]
    lib.dsb_debug_read.restype = ctypes.c_int64
    lib.dsb_condition_launch_count.argtypes = [vp]
    lib.dsb_condition_launch_count.restype = ctypes.c_int64
    lib.dsb_profile_denoise.argtypes = [vp, vp, vp, vp, ci, vp, ctypes.POINTER(cf), ctypes.POINTER(ctypes.c_double),
                                        ctypes.POINTER(ctypes.c_double), ci]
    lib.dsb_profile_denoise.restype = ci
    lib.dsb_profile_timeline.argtypes = [vp, vp, vp, vp, ci, vp, ctypes.POINTER(cf), ctypes.POINTER(cf),
                                         ctypes.POINTER(ci), ci]
    lib.dsb_profile_timeline.restype = ci
    lib.dsb_set_pdl.argtypes = [ci]
    lib.dsb_set_pdl.restype = None
    lib.dsb_get_pdl.argtypes = []
    lib.dsb_get_pdl.restype = ci
    lib.dsb_profile_name.argtypes = [vp, ci]
    lib.dsb_profile_name.restype = ctypes.c_char_p
    lib.dsb_sampler_clamp.argtypes = [vp, ctypes.c_int64, cf, cf, vp]
    lib.dsb_sampler_clamp.restype = ci
    lib.dsb_sampler_dynamic_threshold.argtypes = [vp, ci, ctypes.c_int64, ci, cf, cf, vp]
    lib.dsb_sampler_dynamic_threshold.restype = ci
    lib.dsb_sampler_adaptive_error.argtypes = [vp, vp, vp, ci, ctypes.c_int64, cf, cf, vp, vp]
    lib.dsb_sampler_adaptive_error.restype = ci
    lib.dsb_audio_create.argtypes = [ci, ctypes.POINTER(vp)]
    lib.dsb_audio_create.restype = ci
    lib.dsb_audio_destroy.argtypes = [vp]
    lib.dsb_audio_destroy.restype = None
    lib.dsb_audio_last_error.argtypes = [vp]
    lib.dsb_audio_last_error.restype = ctypes.c_char_p
    lib.dsb_audio_load_weight.argtypes = [vp, ctypes.c_char_p, vp, ctypes.POINTER(ctypes.c_int64), ci]
    lib.dsb_audio_load_weight.restype = ci
    lib.dsb_audio_finalize.argtypes = [vp]
    lib.dsb_audio_finalize.restype = ci
    lib.dsb_audio_forward.argtypes = [vp, vp, vp, ci, vp]
    lib.dsb_audio_forward.restype = ci
    lib.dsb_audio_last_launch_count.argtypes = [vp]
    lib.dsb_audio_last_launch_count.restype = ci
    lib.dsb_vggish_create.argtypes = [ci, ctypes.POINTER(vp)]
    lib.dsb_vggish_create.restype = ci
    lib.dsb_vggish_destroy.argtypes = [vp]
    lib.dsb_vggish_destroy.restype = None
    lib.dsb_vggish_last_error.argtypes = [vp]
    lib.dsb_vggish_last_error.restype = ctypes.c_char_p
    lib.dsb_vggish_load_weight.argtypes = [vp, ctypes.c_char_p, vp, ctypes.POINTER(ctypes.c_int64), ci]
    lib.dsb_vggish_load_weight.restype = ci
    lib.dsb_vggish_finalize.argtypes = [vp]
    lib.dsb_vggish_finalize.restype = ci
    lib.dsb_vggish_forward_feat.argtypes = [vp, vp, vp, ci, vp]
    lib.dsb_vggish_forward_feat.restype = ci
    lib.dsb_vggish_last_launch_count.argtypes = [vp]
    lib.dsb_vggish_last_launch_count.restype = ci
    lib.dsb_mvit_create.argtypes = [ci, ctypes.POINTER(vp)]
    lib.dsb_mvit_create.restype = ci
    lib.dsb_mvit_destroy.argtypes = [vp]
    lib.dsb_mvit_destroy.restype = None
    lib.dsb_mvit_last_error.argtypes = [vp]
    lib.dsb_mvit_last_error.restype = ctypes.c_char_p
    lib.dsb_mvit_load_weight.argtypes = [vp, ctypes.c_char_p, vp, ctypes.POINTER(ctypes.c_int64), ci]
    lib.dsb_mvit_load_weight.restype = ci
    lib.dsb_mvit_finalize.argtypes = [vp]
    lib.dsb_mvit_finalize.restype = ci
    lib.dsb_mvit_forward.argtypes = [vp, vp, ctypes.POINTER(vp), ci, vp]
    lib.dsb_mvit_forward.restype = ci
    lib.dsb_mvit_last_launch_count.argtypes = [vp]
    lib.dsb_mvit_last_launch_count.restype = ci
    lib._dsb_bound = True
    return lib


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class Engine:
    """One handle = one device, one weight set, one workspace sized for ``max_batch`` clips."""

    def __init__(self, max_batch=8, audio_visual=True, device=None):
        if not torch.cuda.is_available():
            raise DsbError("diff_sal_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _bind(_lib.lib())
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.max_batch = int(max_batch)
        self.audio_visual = bool(audio_visual)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            cfg = DsbConfig(self.max_batch, int(self.audio_visual))
            rc = self.lib.dsb_create(ctypes.byref(cfg), ctypes.byref(self._h))
        if rc != 0:
            raise DsbError("dsb_create failed with %d (needs an sm_100 GPU)" % rc)
        self._cond = None
        self.batch = 0

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.dsb_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc < 0 or (rc != 0 and what != "debug"):
            msg = self.lib.dsb_last_error(self._h)
            raise DsbError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else ""))
        return rc

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, state_dict, prefix=""):
        """Loads a reference-keyed SalUNet state_dict (sal_unet.py:146; optional DDP/container
        prefix such as 'module.decoder_net.').  num_batches_tracked entries are ignored."""
        with torch.cuda.device(self.device):
            for key, val in state_dict.items():
                if prefix:
                    if not key.startswith(prefix):
                        continue
                    key = key[len(prefix):]
                if key.endswith("num_batches_tracked"):
                    continue
                t = val.detach().to(dtype=torch.float32).contiguous()
                shape = (ctypes.c_int64 * max(t.dim(), 1))(*t.shape)
                self._check(self.lib.dsb_load_weight(self._h, key.encode(), ctypes.c_void_p(t.data_ptr()), shape, t.dim()),
                            "dsb_load_weight(%s)" % key)
            self._check(self.lib.dsb_finalize_weights(self._h), "dsb_finalize_weights")

    # ------------------------------------------------------------------ condition
    def prepare_condition(self, feat_list, audio=None):
        """Validates the conditioning tensors and returns them as fp32 contiguous tensors on this engine's device
        (the tensors themselves when they already are)."""
        feats = [f.to(device=self.device, dtype=torch.float32).contiguous() for f in feat_list[:3]]
        B = feats[0].shape[0]
        exp = [(768, 8, 7, 12), (384, 8, 14, 24), (192, 8, 28, 48)]
        for f, e in zip(feats, exp):
            if tuple(f.shape[1:]) != e or f.shape[0] != B:
                raise DsbError("feature tensor of shape %s, expected [B,%d,%d,%d,%d]" % ((tuple(f.shape),) + e))
        aud = None
        if audio is not None:
            aud = audio.to(device=self.device, dtype=torch.float32).contiguous()
            if tuple(aud.shape) != (B, 512, 9, 7, 12):
                raise DsbError("audio features of shape %s, expected [%d,512,9,7,12]" % (tuple(aud.shape), B))
        return feats, aud

    def set_condition(self, feat_list, audio=None, prepared=False):
        feats, aud = (feat_list, audio) if prepared else self.prepare_condition(feat_list, audio)
        B = feats[0].shape[0]
        ptrs = (ctypes.c_void_p * 4)(feats[0].data_ptr(), feats[1].data_ptr(), feats[2].data_ptr(), 0)
        with torch.cuda.device(self.device):
            self._check(self.lib.dsb_set_condition(self._h, ptrs, _lib.ptr(aud), B, _stream()), "dsb_set_condition")
        self._cond = (feats, aud)      # keep alive until the enqueued conversion kernels have run
        self.batch = B

    def condition_hash(self, feats, aud=None):
        """64-bit content fingerprint of prepared conditioning tensors (dsb_condition_hash; synchronises the stream)."""
        ptrs = (ctypes.c_void_p * 4)(feats[0].data_ptr(), feats[1].data_ptr(), feats[2].data_ptr(), 0)
        out = ctypes.c_uint64(0)
        with torch.cuda.device(self.device):
            self._check(self.lib.dsb_condition_hash(self._h, ptrs, _lib.ptr(aud), feats[0].shape[0], ctypes.byref(out),
                                                    _stream()), "dsb_condition_hash")
        return int(out.value)

    # ------------------------------------------------------------------ one evaluation
    def denoise(self, x, t, out=None):
        x = x.to(device=self.device, dtype=torch.float32).contiguous()
        B = x.shape[0]
        if tuple(x.shape) != (B, 1, 224, 384):
            raise DsbError("x of shape %s, expected [B,1,224,384]" % (tuple(x.shape),))
        t = torch.as_tensor(t).to(device=self.device, dtype=torch.float32).reshape(-1).contiguous()
        if t.numel() == 1 and B > 1:
            t = t.expand(B).contiguous()
        if t.numel() != B:
            raise DsbError("t has %d entries for a batch of %d" % (t.numel(), B))
        if out is None:
            out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            self._check(self.lib.dsb_denoise(self._h, _lib.ptr(x), _lib.ptr(t), _lib.ptr(out), B, _stream()), "dsb_denoise")
        return out

    def sampler_update(self, coefs, tensors, noise=None, noise_coef=0.0, out=None):
        n = tensors[0].numel()
        if out is None:
            out = torch.empty_like(tensors[0])
        cs = (ctypes.c_float * len(coefs))(*[float(c) for c in coefs])
        ps = (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        with torch.cuda.device(self.device):
            self._check(self.lib.dsb_sampler_update(self._h, cs, ps, len(tensors), _lib.ptr(noise), float(noise_coef),
                                                    _lib.ptr(out), n, _stream()), "dsb_sampler_update")
        return out

    # ------------------------------------------------------------------ whole loop
    def sample(self, ops, x, noise=None, use_graph=True, out_u8=None):
        """Runs a sampler program (list of ('eval', t) / ('axpy', dst, [(src, coef), ...], noise_coef,
        noise_index) / ('clamp', buf, lo, hi) / ('thresh', buf, k, w, max_val) / ('post', buf)) in place on
        x [B,1,224,384] (CUDA, fp32, contiguous, on this engine's device).  ``noise``: [n_slabs,B,1,224,384] on the same
        device.  ``out_u8``: uint8 [B,1,224,384] receiving the min-max-normalised maps of a ('post', buf) op."""
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.device == self.device and x.dtype == torch.float32
                and x.is_contiguous() and x.dim() == 4 and tuple(x.shape[1:]) == (1, 224, 384)):
            raise DsbError("Engine.sample: x must be a contiguous fp32 CUDA tensor [B,1,224,384] on %s, got %s %s on %s"
                           % (self.device, tuple(x.shape), x.dtype, x.device))
        B = x.shape[0]
        if B != self.batch:
            raise DsbError("Engine.sample: batch %d differs from the conditioned batch %d" % (B, self.batch))
        arr = (DsbSamplerOp * len(ops))()
        slabs, post = 0, False
        for i, op in enumerate(ops):
            o = arr[i]
            if op[0] == "eval":                         # ('eval', t[, source buffer])
                o.kind, o.t, o.noise_index = OP_EVAL, float(op[1]), -1
                o.src[0] = int(op[2]) if len(op) > 2 else 0
            elif op[0] == "clamp":                      # ('clamp', buf, lo, hi)
                o.kind, o.dst, o.noise_index = OP_CLAMP, int(op[1]), -1
                o.coef[0], o.coef[1] = float(op[2]), float(op[3])
            elif op[0] == "thresh":                     # ('thresh', buf, k, w, max_val)
                o.kind, o.dst, o.noise_index = OP_DYNTHRESH, int(op[1]), int(op[2])
                o.coef[0], o.coef[1] = float(op[3]), float(op[4])
            elif op[0] == "post":                       # ('post', buf)
                o.kind, o.dst, o.noise_index = OP_POSTPROCESS, int(op[1]), -1
                post = True
            else:
                _, dst, terms, ncoef, nidx = op
                o.kind, o.dst, o.nin = OP_AXPY, int(dst), len(terms)
                for k, (s, c) in enumerate(terms):
                    o.src[k] = int(s)
                    o.coef[k] = float(c)
                o.noise_coef = float(ncoef)
                o.noise_index = int(nidx)
                slabs = max(slabs, int(nidx) + 1)
        if slabs:
            if noise is None:
                raise DsbError("Engine.sample: the program uses %d noise slabs but no noise tensor was given" % slabs)
            if not (noise.is_cuda and noise.device == self.device and noise.dtype == torch.float32 and noise.is_contiguous()
                    and noise.numel() >= slabs * x.numel()):
                raise DsbError("Engine.sample: noise must be a contiguous fp32 CUDA tensor [>=%d,%d,1,224,384] on %s, got "
                               "%s %s on %s" % (slabs, B, self.device, tuple(noise.shape), noise.dtype, noise.device))
        if post:
            if out_u8 is None or not (out_u8.is_cuda and out_u8.device == self.device and out_u8.dtype == torch.uint8
                                      and out_u8.is_contiguous() and out_u8.numel() == x.numel()):
                raise DsbError("Engine.sample: a ('post', buf) op needs out_u8 = contiguous uint8 CUDA tensor [B,1,224,384]")
        desc = DsbSamplerDesc(arr, len(ops), _lib.ptr(noise).value if slabs else None, int(bool(use_graph)),
                              _lib.ptr(out_u8).value if post else None)
        with torch.cuda.device(self.device):
            self._check(self.lib.dsb_sample(self._h, ctypes.byref(desc), _lib.ptr(x), B, _stream()), "dsb_sample")
        return x

    def profile_denoise(self, x, t):
        """[(name, ms, algorithmic_flops, algorithmic_bytes)] for every launch of one evaluation (CUDA events)."""
        x = x.to(device=self.device, dtype=torch.float32).contiguous()
        t = torch.as_tensor(t).to(device=self.device, dtype=torch.float32).reshape(-1).contiguous()
        out = torch.empty_like(x)
        cap = 512
        ms = (ctypes.c_float * cap)()
        fl = (ctypes.c_double * cap)()
        by = (ctypes.c_double * cap)()
        with torch.cuda.device(self.device):
            n = self._check(self.lib.dsb_profile_denoise(self._h, _lib.ptr(x), _lib.ptr(t), _lib.ptr(out), x.shape[0],
                                                         _stream(), ms, fl, by, cap), "debug")
        return [(self.lib.dsb_profile_name(self._h, i).decode(), ms[i], fl[i], by[i]) for i in range(n)]

    def profile_timeline(self, x, t):
        """[(name, start_ms, end_ms, stream)] of one evaluation run on its three streams (CUDA events per launch)."""
        x = x.to(device=self.device, dtype=torch.float32).contiguous()
        t = torch.as_tensor(t).to(device=self.device, dtype=torch.float32).reshape(-1).contiguous()
        out = torch.empty_like(x)
        cap = 512
        a, b, si = (ctypes.c_float * cap)(), (ctypes.c_float * cap)(), (ctypes.c_int * cap)()
        with torch.cuda.device(self.device):
            n = self._check(self.lib.dsb_profile_timeline(self._h, _lib.ptr(x), _lib.ptr(t), _lib.ptr(out), x.shape[0],
                                                          _stream(), a, b, si, cap), "debug")
        return [(self.lib.dsb_profile_name(self._h, i).decode(), a[i], b[i], si[i]) for i in range(n)]

    def condition_launch_count(self):
        return int(self.lib.dsb_condition_launch_count(self._h))

    def last_launch_count(self):
        return int(self.lib.dsb_last_launch_count(self._h))

    def workspace_bytes(self, B=None):
        return int(self.lib.dsb_workspace_bytes(self._h, self.max_batch if B is None else B))

    def debug_read(self, name, numel):
        buf = torch.empty(numel, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            n = self._check(self.lib.dsb_debug_read(self._h, name.encode(), _lib.ptr(buf), numel, _stream()), "debug")
        return buf[:n]


class AudioEngine:
    """Handle of the once-per-clip audio transformer (dsb_audio_*; models/audio_attention.py:93-143)."""

    def __init__(self, max_batch=8, device=None):
        if not torch.cuda.is_available():
            raise DsbError("diff_sal_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _bind(_lib.lib())
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.max_batch = int(max_batch)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.dsb_audio_create(self.max_batch, ctypes.byref(self._h))
        if rc != 0:
            raise DsbError("dsb_audio_create failed with %d (needs an sm_100 GPU)" % rc)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.dsb_audio_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            msg = self.lib.dsb_audio_last_error(self._h)
            raise DsbError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else ""))

    def load_state_dict(self, state_dict, prefix=""):
        """Loads an ``AudioAttnNet.state_dict()`` (optional container prefix such as 'module.spatiotemp_net.')."""
        with torch.cuda.device(self.device):
            for key, val in state_dict.items():
                if prefix:
                    if not key.startswith(prefix):
                        continue
                    key = key[len(prefix):]
                t = val.detach().to(dtype=torch.float32).contiguous()
                shape = (ctypes.c_int64 * max(t.dim(), 1))(*t.shape)
                self._check(self.lib.dsb_audio_load_weight(self._h, key.encode(), ctypes.c_void_p(t.data_ptr()), shape,
                                                           t.dim()), "dsb_audio_load_weight(%s)" % key)
            self._check(self.lib.dsb_audio_finalize(self._h), "dsb_audio_finalize")

    def forward(self, audio):
        """audio [B, 512, 9, 7, 12] -> same shape, fp32 on this engine's device."""
        if audio.dim() != 5 or tuple(audio.shape[1:]) != (512, 9, 7, 12):
            raise DsbError("audio features must be [B, 512, 9, 7, 12], got %s" % (tuple(audio.shape),))
        B = audio.shape[0]
        if B < 1 or B > self.max_batch:
            raise DsbError("batch %d outside [1, %d]" % (B, self.max_batch))
        a = audio.to(device=self.device, dtype=torch.float32).contiguous()
        out = torch.empty_like(a)
        with torch.cuda.device(self.device):
            self._check(self.lib.dsb_audio_forward(self._h, ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                                   B, _stream()), "dsb_audio_forward")
        return out

    @property
    def last_launch_count(self):
        return int(self.lib.dsb_audio_last_launch_count(self._h))


class VggishEngine:
    """Handle of the VGGish feature stack (dsb_vggish_*; models/vggish.py:87-103)."""

    def __init__(self, max_frames=72, device=None):
        if not torch.cuda.is_available():
            raise DsbError("diff_sal_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _bind(_lib.lib())
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.max_frames = int(max_frames)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.dsb_vggish_create(self.max_frames, ctypes.byref(self._h))
        if rc != 0:
            raise DsbError("dsb_vggish_create failed with %d (needs an sm_100 GPU)" % rc)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.dsb_vggish_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            msg = self.lib.dsb_vggish_last_error(self._h)
            raise DsbError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else ""))

    def load_state_dict(self, state_dict, prefix=""):
        with torch.cuda.device(self.device):
            for key, val in state_dict.items():
                if prefix:
                    if not key.startswith(prefix):
                        continue
                    key = key[len(prefix):]
                if not key.startswith("features."):
                    continue                                  # embeddings.*: not on forward_feat's path
                t = val.detach().to(dtype=torch.float32).contiguous()
                shape = (ctypes.c_int64 * max(t.dim(), 1))(*t.shape)
                self._check(self.lib.dsb_vggish_load_weight(self._h, key.encode(), ctypes.c_void_p(t.data_ptr()), shape,
                                                            t.dim()), "dsb_vggish_load_weight(%s)" % key)
            self._check(self.lib.dsb_vggish_finalize(self._h), "dsb_vggish_finalize")

    def forward_feat(self, x):
        """x [frames, 1, 112, 192] -> [frames, 512, 7, 12] fp32 on this engine's device."""
        if x.dim() != 4 or tuple(x.shape[1:]) != (1, 112, 192):
            raise DsbError("audio patches must be [frames, 1, 112, 192], got %s" % (tuple(x.shape),))
        F_ = x.shape[0]
        if F_ < 1 or F_ > self.max_frames:
            raise DsbError("%d frames outside [1, %d]" % (F_, self.max_frames))
        a = x.to(device=self.device, dtype=torch.float32).contiguous()
        out = torch.empty((F_, 512, 7, 12), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.dsb_vggish_forward_feat(self._h, ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                                         F_, _stream()), "dsb_vggish_forward_feat")
        return out

    @property
    def last_launch_count(self):
        return int(self.lib.dsb_vggish_last_launch_count(self._h))


class MvitEngine:
    """Handle of the MViTv2-S video encoder (dsb_mvit_*; models/mvit.py:796-1152)."""

    SHAPES = [(768, 8, 7, 12), (384, 8, 14, 24), (192, 8, 28, 48), (96, 8, 56, 96)]

    def __init__(self, max_batch=2, device=None):
        if not torch.cuda.is_available():
            raise DsbError("diff_sal_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _bind(_lib.lib())
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.max_batch = int(max_batch)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.dsb_mvit_create(self.max_batch, ctypes.byref(self._h))
        if rc != 0:
            raise DsbError("dsb_mvit_create failed with %d (needs an sm_100 GPU)" % rc)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.dsb_mvit_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            msg = self.lib.dsb_mvit_last_error(self._h)
            raise DsbError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else ""))

    def load_state_dict(self, state_dict, prefix=""):
        with torch.cuda.device(self.device):
            for key, val in state_dict.items():
                if prefix:
                    if not key.startswith(prefix):
                        continue
                    key = key[len(prefix):]
                t = val.detach().to(dtype=torch.float32).contiguous()
                shape = (ctypes.c_int64 * max(t.dim(), 1))(*t.shape)
                self._check(self.lib.dsb_mvit_load_weight(self._h, key.encode(), ctypes.c_void_p(t.data_ptr()), shape, t.dim()),
                            "dsb_mvit_load_weight(%s)" % key)
            self._check(self.lib.dsb_mvit_finalize(self._h), "dsb_mvit_finalize")

    def forward(self, video):
        """video [B, 3, 16, 224, 384] (or the loader's raw 4-D view [B*16, 3, 224, 384], mvit.py:1110-1111) -> the four
        feature tensors, coarsest first, fp32 on this engine's device."""
        if video.dim() == 4:
            video = video.reshape(-1, video.shape[-3], 16, video.shape[-2], video.shape[-1])
        if video.dim() != 5 or tuple(video.shape[1:]) != (3, 16, 224, 384):
            raise DsbError("video must be [B, 3, 16, 224, 384], got %s" % (tuple(video.shape),))
        B = video.shape[0]
        if B < 1 or B > self.max_batch:
            raise DsbError("batch %d outside [1, %d]" % (B, self.max_batch))
        v = video.to(device=self.device, dtype=torch.float32).contiguous()
        outs = [torch.empty((B,) + s, dtype=torch.float32, device=self.device) for s in self.SHAPES]
        ptrs = (ctypes.c_void_p * 4)(*[o.data_ptr() for o in outs])
        with torch.cuda.device(self.device):
            self._check(self.lib.dsb_mvit_forward(self._h, ctypes.c_void_p(v.data_ptr()), ptrs, B, _stream()), "dsb_mvit_forward")
        return outs

    @property
    def last_launch_count(self):
        return int(self.lib.dsb_mvit_last_launch_count(self._h))
