"""Shared nn.Module plumbing of the drop-in modules (SalUNetB200, AudioAttnNetB200, VGGishB200).

The reference loads and saves checkpoints on the (DDP-wrapped) PARENT module: ``model.load_state_dict(ckpt, strict=0)``
(model.py:17-22, diffusion_trainer.py:186,728,853) and ``model.state_dict()`` (diffusion_trainer.py:263-280).  nn.Module
implements both by recursing into the children through the ``_load_from_state_dict`` / ``_save_to_state_dict`` hooks with
the child's prefix ('module.decoder_net.' ...), never through the child's own ``load_state_dict`` / ``state_dict``.  The
drop-in modules hold no nn.Parameter (their weights live in the C-ABI handle, repacked for the kernels), so they implement
those two hooks: a reference checkpoint loaded on the unmodified container reaches the engine, and a checkpoint saved from
it carries the reference keys again.
"""
import torch
import torch.nn as nn

from .engine import DsbError


class EngineModule(nn.Module):
    """Subclass contract: ``_spec()`` -> [(key, shape)] in the reference's order, ``_required(key)`` -> bool,
    ``_make_engine()`` -> an engine with ``load_state_dict(dict)`` / ``close()``."""

    def __init__(self):
        super().__init__()
        self._engine = None
        self._sd = None

    # ---- subclass hooks
    def _spec(self):
        raise NotImplementedError

    def _required(self, key):
        return not key.endswith("num_batches_tracked")

    def _make_engine(self):
        raise NotImplementedError

    def _optional(self, key):
        """Keys of the reference module that may be absent even under strict=True (never read on the hot path)."""
        return key.endswith("num_batches_tracked")

    def _weights_changed(self):
        pass

    # ---- nn.Module hooks: what a parent's load_state_dict / state_dict reach
    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        spec = self._spec()
        want = [k for k, _ in spec]
        wanted = set(want)
        have = [k[len(prefix):] for k in state_dict if k.startswith(prefix)]
        if strict:
            unexpected_keys.extend(prefix + k for k in have if k not in wanted)
        present = [k for k in want if prefix + k in state_dict]
        missing = [k for k in want if self._required(k) and prefix + k not in state_dict]
        # tensors the reference module owns but the kernels never read (reported like nn.Module would, fatal only if strict)
        missing_keys.extend(prefix + k for k in want if not self._required(k) and prefix + k not in state_dict
                            and not self._optional(k))
        if not present:
            # a checkpoint without this sub-network (strict=0 loads of a backbone-only file): nothing to build
            missing_keys.extend(prefix + k for k in missing)
            return
        if missing:
            missing_keys.extend(prefix + k for k in missing)
            error_msgs.append("%s: %d of its tensors are missing under '%s' (first: %s); the kernels cannot run on "
                              "partial weights" % (type(self).__name__, len(missing), prefix, missing[0]))
            return
        shapes = dict(spec)
        sd = {}
        for k in present:
            v = state_dict[prefix + k]
            if self._required(k) and tuple(v.shape) != tuple(shapes[k]):
                error_msgs.append("size mismatch for %s%s: checkpoint %s, expected %s" % (prefix, k, tuple(v.shape), tuple(shapes[k])))
                return
            sd[k] = v.detach().to("cpu", torch.float32 if v.is_floating_point() else v.dtype).clone()
        self._sd = sd
        if self._engine is not None:
            self._engine.close()
        self._engine = self._make_engine()
        self._engine.load_state_dict({k: v for k, v in sd.items() if self._required(k)})
        self._weights_changed()

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        if self._sd is None:
            return
        for k, _ in self._spec():
            if k in self._sd:
                destination[prefix + k] = self._sd[k] if keep_vars else self._sd[k].detach()

    # ---- direct calls keep the convenience ``prefix=`` argument and raise DsbError
    def load_state_dict(self, state_dict, strict=True, prefix="", **kwargs):
        if prefix:
            state_dict = {k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)}
        try:
            return super().load_state_dict(state_dict, strict=bool(strict), **kwargs)
        except DsbError:
            raise
        except RuntimeError as e:
            raise DsbError(str(e)) from None

    @property
    def engine(self):
        if self._engine is None:
            raise DsbError("%s has no weights: load a reference state_dict first (directly or through the parent "
                           "module's load_state_dict)" % type(self).__name__)
        return self._engine
