"""The integer execution engine: a calibrated VisionTransformer -> libp2vit_b200.so.

``IntegerEngine`` extracts the calibrated state once, builds one integer plan per ``bit_config``
(plans are cached, so the mixed-precision sweeps of the reference re-use them), uploads the plan's
tensors and binds them to a ``p2v_vit`` handle.  Torch owns every device buffer; the library only
launches kernels on the caller's stream.
"""
import ctypes as C

import numpy as np
import torch

from . import _cabi
from .plan import build_plan, extract_state


def _dev(t, device):
    return t.to(device).contiguous()


class _BoundPlan:
    """A VitPlan uploaded to one device with its C descriptor and library handle."""

    def __init__(self, plan, device):
        self.plan = plan
        self.device = device
        self.keep = []          # device tensors referenced by raw pointers
        lib = _cabi.lib()
        a = plan.arch
        desc = _cabi.VitDesc()
        desc.img_size, desc.patch_size, desc.in_chans = a['img_size'], a['patch_size'], a['in_chans']
        desc.embed_dim, desc.depth, desc.num_heads = a['embed_dim'], a['depth'], a['num_heads']
        desc.hidden_dim, desc.num_classes = a['hidden_dim'], a['num_classes']
        desc.input_scale, desc.input_zp = plan.input_scale, plan.input_zp
        desc.patch_embed = self._linear(plan.patch_embed)
        desc.pe_scale, desc.pe_zp = plan.pe_scale, plan.pe_zp
        desc.embed_scale, desc.embed_zp = plan.embed_scale, plan.embed_zp
        desc.cls_value = self._p(plan.cls_value)
        desc.pos_value = self._p(plan.pos_value)
        desc.embed_out_scale = self._p(plan.embed_out_scale)
        self.blocks = (_cabi.BlockDesc * a['depth'])()
        for i, b in enumerate(plan.blocks):
            bd = self.blocks[i]
            bd.norm1, bd.norm2 = self._ln(b.norm1), self._ln(b.norm2)
            bd.qkv, bd.proj = self._linear(b.qkv), self._linear(b.proj)
            bd.fc1, bd.fc2 = self._linear(b.fc1), self._linear(b.fc2)
            bd.attn = self._attn(b.attn)
        desc.blocks = C.cast(self.blocks, C.POINTER(_cabi.BlockDesc))
        desc.norm = self._ln(plan.norm)
        desc.head = self._linear(plan.head)
        self.desc = desc
        handle = C.c_void_p()
        _cabi.check(lib.p2v_vit_create(C.byref(desc), device.index or 0, C.byref(handle)))
        self.handle = handle
        self.launches = lib.p2v_vit_launches_per_forward(handle)
        self._buffers = {}

    def _p(self, t):
        if t is None:
            return None
        d = _dev(t, self.device)
        self.keep.append(d)
        return d.data_ptr()

    def _linear(self, lp):
        d = _cabi.LinearDesc()
        if lp.w is not None:
            d.w = self._p(lp.w)
            d.n, d.k = lp.w.shape
        else:   # int4-packed layer of a serialised plan: nibbles are expanded to int8 codes on the device
            packed = _dev(lp.w4, self.device)
            n, k = lp.w4.shape[0], lp.w4.shape[1] * 2
            codes = torch.empty((n, k), dtype=torch.int8, device=self.device)
            _cabi.check(_cabi.lib().p2v_unpack_int4(packed.data_ptr(), codes.data_ptr(), packed.numel(),
                                                    _cabi.current_stream(self.device)))
            torch.cuda.current_stream(self.device).synchronize()
            self.keep.append(codes)
            d.w = codes.data_ptr()
            d.n, d.k = n, k
        e = d.epi
        e.acc_scale, e.bias = self._p(lp.acc_scale), self._p(lp.bias)
        e.out_scale, e.out_rscale = self._p(lp.out_scale), self._p(lp.out_rscale)
        e.res_scale, e.out2_scale = self._p(lp.res_scale), self._p(lp.out2_scale)
        e.out_zp, e.flags = lp.out_zp, lp.flags
        return d

    def _ln(self, p):
        d = _cabi.LayerNorm()
        d.in_mask, d.gamma, d.beta = self._p(p.in_mask), self._p(p.gamma), self._p(p.beta)
        d.ln_out_scale, d.ln_out_rscale = self._p(p.ln_out_scale), self._p(p.ln_out_rscale)
        d.post_mul, d.post_div1 = self._p(p.post_mul), self._p(p.post_div1)
        d.post_div2, d.post_zp, d.in_scale1, d.pot = p.post_div2, p.post_zp, p.in_scale1, p.pot
        return d

    def _attn(self, p):
        d = _cabi.Attention()
        d.score_mul, d.score_zp, d.out_mul, d.out_zp = p.score_mul, p.score_zp, p.out_mul, p.out_zp
        d.softmax_levels = p.levels
        d.in_zp = getattr(p, 'in_zp', 0.0)
        d.lut_sig_bits = p.lut_sig_bits
        d.exp_lut = self._p(p.exp_lut)
        return d

    def buffers(self, batch, slot=0):
        """(workspace, logits, logit codes) for a batch size; allocated once so graph replays stay valid.
        `slot` > 0 gives additional logits/code buffers over the same workspace (pipelined host I/O)."""
        if slot:
            key = ('slot', batch, slot)
            if key not in self._buffers:
                ws, ws_ptr, _, _ = self.buffers(batch)
                logits = torch.empty(batch, self.plan.arch['num_classes'], dtype=torch.float32, device=self.device)
                codes = torch.empty(batch, self.plan.arch['num_classes'], dtype=torch.int8, device=self.device)
                self._buffers[key] = (ws, ws_ptr, logits, codes)
            return self._buffers[key]
        if batch not in self._buffers:
            nbytes = _cabi.lib().p2v_vit_workspace_bytes(self.handle, batch)
            ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
            off = (-ws.data_ptr()) % 1024
            logits = torch.empty(batch, self.plan.arch['num_classes'], dtype=torch.float32, device=self.device)
            codes = torch.empty(batch, self.plan.arch['num_classes'], dtype=torch.int8, device=self.device)
            self._buffers[batch] = (ws, ws.data_ptr() + off, logits, codes)
        return self._buffers[batch]

    def close(self):
        if self.handle:
            _cabi.lib().p2v_vit_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class IntegerEngine:
    """Quantized forward of one calibrated model on one CUDA device."""

    def __init__(self, model=None, state=None, device=None, plans=()):
        """`plans`: ready-made VitPlans (e.g. `plan.load_plan(path)`), used for their bit_config without any
        calibrated model or state."""
        if state is None and model is not None:
            state = extract_state(model)
        self.state = state
        if device is None:
            device = next(model.parameters()).device if model is not None else torch.device('cuda', 0)
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            if torch.cuda.is_available():
                self.device = torch.device('cuda', torch.cuda.current_device())
            else:
                raise RuntimeError('the quantized forward runs only in the sm_100a kernels of libp2vit_b200.so; '
                                   'no CUDA device is available and there is no CPU fallback')
        if self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        self._plans = {}
        for plan in plans:
            self._plans[tuple(int(b) for b in plan.bit_config)] = _BoundPlan(plan, self.device)

    max_plans = 8   # device-resident plans kept per engine (a mixed-precision search walks hundreds of configs)

    def bound(self, bit_config):
        """The device-resident plan of a bit_config: re-selected from the calibrated state (no re-calibration),
        cached, least recently used evicted beyond `max_plans`."""
        key = tuple(int(b) for b in bit_config)
        if key in self._plans:
            self._plans[key] = self._plans.pop(key)          # most recently used last
            return self._plans[key]
        if self.state is None:
            raise KeyError('no plan for bit_config %s and no calibrated state to build one from' % (key,))
        while len(self._plans) >= self.max_plans:
            torch.cuda.synchronize(self.device)              # nothing in flight may still read the evicted buffers
            self._plans.pop(next(iter(self._plans))).close()
        self._plans[key] = _BoundPlan(build_plan(self.state, key), self.device)
        return self._plans[key]

    def _check_input(self, x, arch):
        if x.dim() != 4 or x.shape[1] != arch['in_chans'] or x.shape[2] != arch['img_size'] or x.shape[3] != arch['img_size']:
            raise AssertionError("Input image size (%d*%d) doesn't match model (%d*%d)." %
                                 (x.shape[-2], x.shape[-1], arch['img_size'], arch['img_size']))

    def forward_into(self, x, bit_config, use_graph=True, slot=0):
        """Launch the forward on the current stream; returns the engine-owned logits buffer (of `slot`)."""
        bp = self.bound(bit_config)
        self._check_input(x, bp.plan.arch)
        if not x.is_cuda:
            raise RuntimeError('IntegerEngine.forward_into expects a CUDA tensor')
        x = x.contiguous().float()
        b = x.shape[0]
        _, ws, logits, codes = bp.buffers(b, slot)
        _cabi.check(_cabi.lib().p2v_vit_forward(bp.handle, x.data_ptr(), logits.data_ptr(), codes.data_ptr(), b, ws,
                                                None, 1 if use_graph else 0, _cabi.current_stream(self.device)))
        self._last_input = x  # keep the (possibly re-laid-out) input alive until the stream consumed it
        return logits

    def forward_into_u8(self, x_u8, bit_config, mean, std, use_graph=True, slot=0):
        """The forward from 8-bit pixels [B, C, H, W] (CUDA uint8) with the loader's normalisation constants: the device
        evaluates (pixel / 255 - mean[c]) / std[c] op for op, so the logits equal `forward_into` on the normalised fp32
        tensor, for a quarter of the input traffic.  Returns the engine-owned logits buffer (of `slot`)."""
        bp = self.bound(bit_config)
        self._check_input(x_u8, bp.plan.arch)
        if not x_u8.is_cuda or x_u8.dtype != torch.uint8:
            raise RuntimeError('IntegerEngine.forward_into_u8 expects a CUDA uint8 tensor')
        x_u8 = x_u8.contiguous()
        b, c = x_u8.shape[0], x_u8.shape[1]
        m = (C.c_float * c)(*[float(v) for v in mean])
        sd = (C.c_float * c)(*[float(v) for v in std])
        _, ws, logits, codes = bp.buffers(b, slot)
        _cabi.check(_cabi.lib().p2v_vit_forward_u8(bp.handle, x_u8.data_ptr(), m, sd, logits.data_ptr(), codes.data_ptr(),
                                                   b, ws, 1 if use_graph else 0, _cabi.current_stream(self.device)))
        self._last_input = x_u8
        return logits

    def forward(self, x, bit_config):
        """Drop-in: fp32 logits [B, classes] on x's device (host inputs are copied to the GPU and back)."""
        on_host = not x.is_cuda
        bp = self.bound(bit_config)
        self._check_input(x, bp.plan.arch)
        # the caller's tensor lives at a new address on most calls; the captured graph is keyed on its input address,
        # so the batch is copied into an engine-owned staging buffer (one per batch size) and the graph is replayed,
        # instead of paying a stream capture + instantiation (milliseconds on the host) per call
        key = ('xstage', x.shape[0])
        if key not in bp._buffers:
            bp._buffers[key] = torch.empty(tuple(x.shape), dtype=torch.float32, device=self.device)
        stage = bp._buffers[key]
        stage.copy_(x, non_blocking=True)
        out = self.forward_into(stage, bit_config).clone()
        return out.cpu() if on_host else out

    def forward_dump(self, x, bit_config):
        """Forward that also returns every intermediate integer code tensor (parity tests, hooks)."""
        bp = self.bound(bit_config)
        self._check_input(x, bp.plan.arch)
        lib = _cabi.lib()
        x = x.to(self.device).contiguous().float()
        b = x.shape[0]
        _, ws, logits, codes = bp.buffers(b)
        nbytes = lib.p2v_vit_dump_bytes(bp.handle, b)
        dump = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        _cabi.check(lib.p2v_vit_forward(bp.handle, x.data_ptr(), logits.data_ptr(), codes.data_ptr(), b, ws,
                                        dump.data_ptr(), 0, _cabi.current_stream(self.device)))
        torch.cuda.synchronize(self.device)
        host = dump.cpu().numpy()
        out = {}
        name, off, size, elem = C.c_char_p(), C.c_int64(), C.c_int64(), C.c_int32()
        count = lib.p2v_vit_dump_layout(bp.handle, b, 0, C.byref(name), C.byref(off), C.byref(size), C.byref(elem))
        for i in range(count):
            lib.p2v_vit_dump_layout(bp.handle, b, i, C.byref(name), C.byref(off), C.byref(size), C.byref(elem))
            raw = host[off.value:off.value + size.value]
            key = name.value.decode()
            if elem.value == 4:
                out[key] = raw.view(np.int32).copy()
            elif key.startswith('softmax/'):
                out[key] = raw.view(np.uint8).copy()
            else:
                out[key] = raw.view(np.int8).copy()
        return logits.clone(), out

    def forward_host_pipelined(self, batches_host, logits_host, bit_config, mean=None, std=None, after_forward=None):
        """Serving loop over HOST batches: the pinned-memory H2D copy of batch i+1 and the D2H copy of the logits
        of batch i-1 overlap the forward of batch i (two device input buffers, two logits buffers, one copy
        stream, CUDA events).  `batches_host[i]` -> `logits_host[i]`; returns after everything has landed.
        uint8 batches (raw pixels) need the loader's `mean` / `std` and take the 8-bit entry (`forward_into_u8`).
        `after_forward(logits, slot)` (optional) runs on the compute stream right after a forward and returns the
        tensor to send to the host instead of the local logits (e.g. their NCCL all-gather, slot in {0, 1})."""
        bp = self.bound(bit_config)
        u8 = batches_host[0].dtype == torch.uint8
        if u8 and (mean is None or std is None):
            raise ValueError('uint8 batches need the normalisation constants mean and std')
        comp = torch.cuda.current_stream(self.device)
        if not hasattr(self, '_io_streams'):
            self._io_streams = (torch.cuda.Stream(self.device), torch.cuda.Stream(self.device))
        h2d, d2h = self._io_streams        # separate queues: an upload never waits behind a download
        b = batches_host[0].shape[0]
        key = ('xpipe', b, u8)
        if key not in bp._buffers:
            bp._buffers[key] = ([torch.empty(batches_host[0].shape, dtype=batches_host[0].dtype, device=self.device) for _ in range(2)],
                                [[torch.cuda.Event() for _ in range(2)] for _ in range(3)])
        xdev, (ev_in, ev_done, ev_out) = bp._buffers[key]
        h2d.wait_stream(comp)
        d2h.wait_stream(comp)
        for i, xh in enumerate(batches_host):
            slot = i & 1
            with torch.cuda.stream(h2d):
                if i >= 2:
                    h2d.wait_event(ev_done[slot])           # the forward that read this input buffer has finished
                xdev[slot].copy_(xh, non_blocking=True)
                ev_in[slot].record(h2d)
            comp.wait_event(ev_in[slot])
            if i >= 2:
                comp.wait_event(ev_out[slot])               # the logits of batch i-2 have left their buffer
            if u8:
                logits = self.forward_into_u8(xdev[slot], bit_config, mean, std, slot=slot + 1)
            else:
                logits = self.forward_into(xdev[slot], bit_config, slot=slot + 1)
            if after_forward is not None:
                logits = after_forward(logits, slot)
            ev_done[slot].record(comp)
            with torch.cuda.stream(d2h):
                d2h.wait_event(ev_done[slot])
                logits_host[i].copy_(logits, non_blocking=True)
                ev_out[slot].record(d2h)
        h2d.synchronize()
        d2h.synchronize()
        comp.synchronize()
        return logits_host

    def forward_host(self, x_host, logits_host, bit_config):
        """End-to-end call on HOST buffers (pinned for async copies): H2D, forward, D2H, sync."""
        bp = self.bound(bit_config)
        b = x_host.shape[0]
        _, ws, logits, _ = bp.buffers(b)
        key = ('xdev', b)
        if key not in bp._buffers:
            bp._buffers[key] = torch.empty(x_host.shape, dtype=torch.float32, device=self.device)
        xdev = bp._buffers[key]
        _cabi.check(_cabi.lib().p2v_vit_forward_host(bp.handle, x_host.data_ptr(), logits_host.data_ptr(), b, ws,
                                                     xdev.data_ptr(), logits.data_ptr(), _cabi.current_stream(self.device)))
        return logits_host
