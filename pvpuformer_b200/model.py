"""Drop-in host side of the per-click forward: the reference's nn.Module surface over the C ABI.

Mirrors `VitMultiGaussianVector_ed_Model` (reference isegm/model/is_vpu_model.py:140-438):
same constructor arguments, same state_dict keys (so `load_state_dict(reference.state_dict(),
strict=True)` works), same attributes the predictor plumbing reads (`with_prev_mask`,
`backbone.pos_embed`, `backbone.patch_embed.{grid_size,num_patches,patch_size}`, `_config`) and the
same `forward(image, points, prompts, as_prompt_type, edloss, pclout) -> {'instances',
'instances_aux'}`.  All arithmetic happens in libvpuformer_b200.so (hand-written sm_100a CUDA);
torch only owns memory and streams.  There is no fallback: without the extension or on a
non-sm_100 device the forward raises.
"""
import copy
import ctypes
import functools
import inspect
import random

import numpy as np
import torch
import torch.nn as nn

from . import host_prompts, lib as L, ops
from .config import VPUConfig
from .packing import pack_weights
from .weights import param_spec

_ARCH_BY_DIM = {768: "vit_base", 1024: "vit_large", 1280: "vit_huge"}


class _Holder(nn.Module):
    """Parameter container; nested holders reproduce the reference's dotted state_dict keys."""


def _register(root, dotted, tensor, kind):
    parts = dotted.split(".")
    m = root
    for p in parts[:-1]:
        if p not in m._modules:
            m.add_module(p, _Holder())
        m = m._modules[p]
    if kind == "buffer":
        m.register_buffer(parts[-1], tensor)
    else:
        m.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


REFERENCE_CLASS = "isegm.model.is_vpu_model.VitMultiGaussianVector_ed_Model"


def _record_config(init):
    """What the reference's @serialize decorator records (utils/serialization.py:7-41): `_config` = {'class': dotted name of
    the reference class, 'params': {name: {'type': 'builtin' | 'class', 'value', 'specified'}}} for every constructor
    argument, so that `save_checkpoint` (utils/misc.py:31-33) writes a checkpoint either implementation loads."""
    names = list(inspect.signature(init).parameters)[1:]
    defaults = {k: v.default for k, v in inspect.signature(init).parameters.items() if v.default is not inspect.Parameter.empty}

    @functools.wraps(init)
    def new_init(self, *args, **kwargs):
        params = copy.deepcopy(kwargs)
        params.update(zip(names, args))
        specified = set(params)
        for k, v in defaults.items():
            params.setdefault(k, v)
        cfg = {"class": REFERENCE_CLASS, "params": {}}
        for k, v in params.items():
            is_class = inspect.isclass(v)
            cfg["params"][k] = {"type": "class" if is_class else "builtin",
                                "value": (v.__module__ + "." + v.__qualname__) if is_class else v, "specified": k in specified}
        self._config = cfg
        init(self, *args, **kwargs)
    return new_init


class VitMultiGaussianVector_ed_Model(nn.Module):
    @_record_config
    def __init__(self, num_max_points=24, backbone_params={}, neck_params={}, head_params={}, random_split=False,
                 residual=False, with_aux_output=False, norm_radius=5, use_disks=False, cpu_dist_maps=False,
                 use_rgb_conv=False, use_leaky_relu=False, with_prev_mask=False,
                 norm_mean_std=([.485, .456, .406], [.229, .224, .225])):
        super().__init__()
        if random_split:
            raise NotImplementedError("random_split=True (token shuffle, models_vit.py:266-272) is outside the B200 path")
        if not (use_disks and with_prev_mask) or cpu_dist_maps or use_rgb_conv:
            raise NotImplementedError("the B200 path implements the shipped VPU config only: use_disks=True, "
                                      "with_prev_mask=True, cpu_dist_maps=False, use_rgb_conv=False "
                                      "(models/iSegNet/vpu_base448_cocolvis.py:46-56)")
        bp = dict(backbone_params)
        img = bp.get("img_size", (448, 448))
        patch = bp.get("patch_size", (16, 16))
        if img[0] != img[1] or patch[0] != patch[1]:
            raise NotImplementedError("square images / patches only")
        C = bp.get("embed_dim", 768)
        hp = dict(head_params)
        self.cfg = VPUConfig(arch=_ARCH_BY_DIM.get(C, "custom"), img_size=img[0], patch=patch[0], embed_dim=C,
                             depth=bp.get("depth", 12), num_heads=bp.get("num_heads", 12),
                             num_max_points=num_max_points, head_channels=hp.get("channels", 256),
                             out_dims=tuple(neck_params.get("out_dims", [128, 256, 512, 1024])),
                             norm_radius=norm_radius, norm_mean=tuple(norm_mean_std[0]), norm_std=tuple(norm_mean_std[1]))
        if hp.get("upsample", "x1") != "x1" or hp.get("align_corners", False):
            raise NotImplementedError("head upsample='x1', align_corners=False only")
        self.with_aux_output = with_aux_output
        self.with_prev_mask = with_prev_mask
        self.with_points = False
        self.num_max_points = num_max_points
        self.random_split = random_split
        self.residual = residual
        self.want_aux = True          # set False to skip the 48-channel aux upsample (NoBRS only reads 'instances')
        # click-prompt forwards of at most this many samples (the interactive NoBRS shape: one click + flip TTA = 2) are
        # captured once per (batch, clicks, aux) into a CUDA graph and replayed: ~185 launches become one (0 disables)
        self.graph_max_batch = 8
        self._graphs = {}
        for key, (shape, kind) in param_spec(self.cfg).items():
            _register(self, key, torch.zeros(shape), kind)
        pe = self.backbone.patch_embed
        pe.grid_size = (self.cfg.grid, self.cfg.grid)
        pe.num_patches = self.cfg.num_tokens
        pe.patch_size = (self.cfg.patch, self.cfg.patch)
        pe.img_size = (self.cfg.img_size, self.cfg.img_size)
        self._handle = None
        self._packed = None
        self._ws = {}
        self.register_load_state_dict_post_hook(lambda m, k: m._invalidate())

    # ---- lifecycle -----------------------------------------------------------------------
    def _invalidate(self):
        self._packed = None

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._invalidate()
        return r

    def __del__(self):
        try:
            if self._handle is not None:
                L.load().vpu_destroy(self._handle)
        except Exception:
            pass

    def _ensure_ready(self, device):
        lib = L.load()
        if self._handle is None:
            c = self.cfg
            d = L.VpuDims(c.img_size, c.patch, c.embed_dim, c.depth, c.num_heads, c.num_max_points, c.dma_depth,
                          c.dma_heads, c.dma_mlp_dim, c.ppue_ffn_dim, c.head_channels, (ctypes.c_int32 * 4)(*c.out_dims),
                          float(c.norm_radius))
            h = ctypes.c_void_p()
            L.check(lib.vpu_create(ctypes.byref(h), ctypes.byref(d)))
            self._handle = h
            # the reference's float32 numpy formula for the click taps (ops.py:51-61)
            k = np.arange(0, 19, 1, np.float32)
            t = np.exp(-((k - 9) ** 2) / (2 * 3 ** 2)).astype(np.float32)
            t[9] += 1
            L.check(lib.vpu_set_click_table(h, t.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 19))
        if self._packed is None or self._packed[0] != device:
            packed, scalars = pack_weights(self.state_dict(), self.cfg, device)
            for key, t in packed.items():
                shape = (ctypes.c_int64 * t.dim())(*t.shape)
                dt = L.VPU_BF16 if t.dtype == torch.bfloat16 else L.VPU_F32
                L.check(lib.vpu_bind_weight(self._handle, key.encode(), L.ptr(t), dt, shape, t.dim()))
            for key, v in scalars.items():
                L.check(lib.vpu_set_scalar(self._handle, key.encode(), v))
            L.check(lib.vpu_finalize(self._handle))
            self._packed = (device, packed)
            self._ws = {}
            self._graphs = {}

    def workspace(self, B, device):
        """One buffer, sized for the largest batch seen so far and reused by every smaller one (the static plan of
        vpu_workspace_bytes(B) always starts at offset 0): a NoC loop whose active set shrinks click by click never
        re-allocates.  Growing replaces the buffer; the old one is returned to torch's caching allocator, which keeps it
        alive until the work already queued on it has run."""
        need = L.load().vpu_workspace_bytes(self._handle, B)
        ws = self._ws.get("buf")
        if ws is None or ws.numel() < need + 1024 or ws.device != device:
            ws = torch.empty(need + 1024, dtype=torch.uint8, device=device)
            self._ws = {"buf": ws}
        off = (-ws.data_ptr()) % 1024
        return ws[off:]

    def _serialize_streams(self):
        """The workspace is shared by every forward of this module: a forward issued on another stream than the previous
        one first waits for it (device-side event wait, no host sync), so two pipelines over one model cannot race on the
        activations."""
        cur = torch.cuda.current_stream()
        last = getattr(self, "_last_use", None)
        if last is not None and last[0] != cur.cuda_stream:
            cur.wait_event(last[1])
        return cur

    def _mark_use(self, stream):
        last = getattr(self, "_last_use", None)
        ev = last[1] if last is not None else torch.cuda.Event()
        ev.record(stream)
        self._last_use = (stream.cuda_stream, ev)

    def tap(self, name, B, dtype, shape):
        """View of a named intermediate of the LAST forward with batch B (parity tests)."""
        off, nbytes = ctypes.c_size_t(), ctypes.c_size_t()
        L.check(L.load().vpu_workspace_lookup(self._handle, B, name.encode(), ctypes.byref(off), ctypes.byref(nbytes)))
        ws = self.workspace(B, self._packed[0])
        return ws[off.value:off.value + nbytes.value].view(dtype)[:int(np.prod(shape))].view(*shape)

    # ---- prompts -------------------------------------------------------------------------
    def _prompt_struct(self, points, prompts, as_prompt_type, B, device, keep):
        if points is None:
            raise ValueError("points is required")
        if points.dim() != 3 or points.shape[0] != B or points.shape[2] != 3 or points.shape[1] % 2:
            raise ValueError("points must be [B, 2n, 3], got %s" % (tuple(points.shape),))
        n = points.shape[1] // 2
        if n < 1 or n > self.num_max_points:
            raise ValueError("n=%d points per half outside [1, %d] (reference pads to num_max_points, "
                             "is_vpu_model.py:218-228)" % (n, self.num_max_points))
        pts = points.to(device=device, dtype=torch.float64).contiguous()
        keep.append(pts)
        pr = L.VpuPrompts()
        pr.points = pts.data_ptr()
        pr.n = n
        pr.type = int(as_prompt_type)
        if as_prompt_type not in (0, 1, 2):
            raise ValueError("as_prompt_type must be 0, 1 or 2")
        if as_prompt_type != 0:
            if prompts is None:
                raise ValueError("prompts=(points, boxes, [scribbles, rects]) is required for as_prompt_type != 0")
            p_pts, boxes, (scribbles, rects) = prompts
            pp = p_pts.to(device=device, dtype=torch.float64).contiguous()      # is_vpu_model.py:396-397
            keep.append(pp)
            pr.ppue_points = pp.data_ptr()
            pr.n_ppue = pp.shape[1] // 2
            size = self.cfg.img_size
            if as_prompt_type == 1:
                bx = boxes.to(device=device, dtype=torch.int32).contiguous()
                keep.append(bx)
                pr.boxes = bx.data_ptr()
                em = ops.raster_prompts(1, bx, None, n, B, size)        # cv2.rectangle(..., 3) of draw_box, on the device
            else:
                sel = np.stack([host_prompts.scribble_select(np.asarray(scribbles)[b][0], np.asarray(rects)[b][0],
                                                             self.cfg.img_size, random) for b in range(B)])
                slots = host_prompts.scribble_slots(p_pts.detach().cpu().numpy(), pr.n_ppue)
                sel_t = torch.from_numpy(sel).to(device)
                slot_t = torch.from_numpy(slots).to(device)
                keep += [sel_t, slot_t]
                pr.scrib_sel = sel_t.data_ptr()
                pr.scrib_slot = slot_t.data_ptr()
                curve = np.stack([np.asarray(scribbles[b][0])[:, :2] for b in range(B)]).astype(np.int32)     # is_model.py:128 (x, y)
                cv = torch.from_numpy(curve).to(device)                   # cv2.polylines(..., 3) of draw_scribble, on the device
                keep.append(cv)
                em = ops.raster_prompts(2, None, cv, n, B, size)
            keep.append(em)
            pr.extra_mask = em.data_ptr()
        return pr

    def _check_image(self, image):
        s = self.cfg.img_size
        if not image.is_cuda:
            raise L.VpuError("the B200 path runs on a CUDA device only (no CPU fallback); got a %s tensor" % image.device)
        if image.dim() != 4 or image.shape[1] != 4 or image.shape[2] != s or image.shape[3] != s:
            raise ValueError("image must be [B, 4, %d, %d] (RGB + previous mask), got %s" % (s, s, tuple(image.shape)))
        return image.to(torch.float32).contiguous()

    # ---- the operator surface ------------------------------------------------------------
    @torch.no_grad()
    def forward(self, image, points=None, prompts=None, as_prompt_type=0, edloss=True, pclout=False):
        if not edloss:
            raise NotImplementedError("edloss=False is not a working path in the reference either (is_vpu_model.py:411-419)")
        image = self._check_image(image)
        device, B = image.device, image.shape[0]
        self._ensure_ready(device)
        keep = []
        pr = self._prompt_struct(points, prompts, as_prompt_type, B, device, keep)
        s = self.cfg.img_size
        want_aux = self.with_aux_output and self.want_aux
        if as_prompt_type == 0 and 0 < B <= self.graph_max_batch and not torch.cuda.is_current_stream_capturing():
            return self._forward_graph(image, keep[0], B, want_aux, device)
        inst = torch.empty(B, 1, s, s, dtype=torch.float32, device=device)
        aux = torch.empty(B, self.cfg.num_queries, s, s, dtype=torch.float32, device=device) if want_aux else None
        ws = self.workspace(B, device)
        stream = self._serialize_streams()
        L.check(L.load().vpu_forward(self._handle, L.ptr(image), ctypes.byref(pr), B, L.ptr(inst), L.ptr(aux),
                                     L.ptr(ws), ws.numel(), L.current_stream()))
        self._mark_use(stream)
        self._keepalive = (keep, image)     # inputs must outlive the asynchronous kernels
        return {"instances": inst, "instances_aux": aux}

    def _forward_graph(self, image, pts, B, want_aux, device):
        """Small click-prompt batches: vpu_forward never allocates or synchronises and enqueues on one stream only, so the
        whole forward is captured once into a CUDA graph over static input / output buffers and replayed per call (copy the
        inputs in, replay, clone the outputs out).  The click rows are re-laid out into the full 2 x num_max_points slots
        (unused slots = (-1, -1, -1), exactly what the reference pads with, is_vpu_model.py:218-228), so one graph per
        (batch, aux) serves a session whose click count grows with every click.  The graph holds the workspace address: it is
        re-captured if a larger batch has replaced the workspace since."""
        s, n, nm = self.cfg.img_size, pts.shape[1] // 2, self.num_max_points
        ws = self.workspace(B, device)
        key = (B, want_aux, device)
        g = self._graphs.get(key)
        stream = self._serialize_streams()

        def load_inputs():
            g["image"].copy_(image)
            g["pts"].fill_(-1.0)
            g["pts"][:, :n].copy_(pts[:, :n])
            g["pts"][:, nm:nm + n].copy_(pts[:, n:])
        if g is None or g["ws_ptr"] != ws.data_ptr():
            g = {"ws_ptr": ws.data_ptr(), "image": torch.empty_like(image),
                 "pts": torch.empty(B, 2 * nm, 3, dtype=torch.float64, device=device),
                 "inst": torch.empty(B, 1, s, s, dtype=torch.float32, device=device),
                 "aux": torch.empty(B, self.cfg.num_queries, s, s, dtype=torch.float32, device=device) if want_aux else None}
            pr = L.VpuPrompts()
            pr.points = g["pts"].data_ptr()
            pr.n = nm
            pr.type = 0
            load_inputs()

            def enqueue():
                L.check(L.load().vpu_forward(self._handle, L.ptr(g["image"]), ctypes.byref(pr), B, L.ptr(g["inst"]), L.ptr(g["aux"]),
                                             L.ptr(ws), ws.numel(), L.current_stream()))
            enqueue()                        # once outside the capture: tensor maps, function attributes, lazy module loading
            torch.cuda.current_stream().synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                enqueue()
            g["graph"] = graph
            self._graphs[key] = g
        load_inputs()
        g["graph"].replay()
        self._mark_use(stream)
        return {"instances": g["inst"].clone(), "instances_aux": g["aux"].clone() if want_aux else None}

    @torch.no_grad()
    def ppue(self, points, prompts=None, as_prompt_type=0):
        """PPuE rows [B, 48, 899] fp32 (reference _guassinvector_{click,box,scribble})."""
        device = points.device
        if not points.is_cuda:
            raise L.VpuError("CUDA tensors only")
        self._ensure_ready(device)
        B = points.shape[0]
        keep = []
        pr = self._prompt_struct(points, prompts, as_prompt_type, B, device, keep)
        out = torch.empty(B, self.cfg.num_queries, self.cfg.ppue_dim, dtype=torch.float32, device=device)
        L.check(L.load().vpu_ppue(self._handle, ctypes.byref(pr), B, L.ptr(out), L.current_stream()))
        torch.cuda.current_stream().synchronize()
        return out

    @torch.no_grad()
    def coord_features(self, image, points, prompts=None, as_prompt_type=0):
        """cat(prev_mask, disks | raster) [B, 3, H, W] fp32 (reference get_coord_features_with_prompt)."""
        image = self._check_image(image)
        device, B = image.device, image.shape[0]
        self._ensure_ready(device)
        keep = []
        pr = self._prompt_struct(points, prompts, as_prompt_type, B, device, keep)
        out = torch.empty(B, 3, self.cfg.img_size, self.cfg.img_size, dtype=torch.float32, device=device)
        L.check(L.load().vpu_coord_features(self._handle, L.ptr(image), ctypes.byref(pr), B, L.ptr(out), L.current_stream()))
        torch.cuda.current_stream().synchronize()
        return out


def build_model(arch="vit_base", img_size=448, state_dict=None, device=None, with_aux_output=True):
    """The shipped VPU configuration (reference models/iSegNet/vpu_base448_cocolvis.py:17-56) for B/L/H."""
    from .config import make_config
    c = make_config(arch, img_size=img_size)
    m = VitMultiGaussianVector_ed_Model(
        use_disks=True, norm_radius=5, with_prev_mask=True,
        backbone_params=dict(img_size=(img_size, img_size), patch_size=(c.patch, c.patch), in_chans=3, embed_dim=c.embed_dim,
                             depth=c.depth, num_heads=c.num_heads, mlp_ratio=4, qkv_bias=True),
        neck_params=dict(in_dim=c.embed_dim, out_dims=[128, 256, 512, 1024], img_size=(img_size, img_size)),
        head_params=dict(in_channels=[128, 256, 512, 1024], in_index=[0, 1, 2, 3], dropout_ratio=0.1, num_classes=1,
                         loss_decode=None, align_corners=False, upsample="x1", ed_loss=True, channels=256),
        random_split=False, residual=True, with_aux_output=with_aux_output)
    if state_dict is not None:
        m.load_state_dict(state_dict, strict=True)
    if device is not None:
        m.to(device)
    return m.eval()
