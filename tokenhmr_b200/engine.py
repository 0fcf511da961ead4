"""TokenHMREngine — drop-in for the reference's TokenHMR.forward(batch) -> dict surface.

Reference boundary: tokenhmr/lib/models/tokenhmr.py:330-338 (forward) -> :135-188
(forward_step), built by load_tokenhmr (tokenhmr/lib/models/__init__.py:3-26) and called as
`model(batch)` under torch.no_grad() by demo.py:77-78, eval.py:146-147 and track.py:39.

    model = TokenHMREngine(cfg, state_dict, smpl_buffers, device='cuda:0')
    out = model({'img': img})          # img (B,3,256,256) fp32, any device; extra keys ignored
    out['pred_vertices'], out['pred_keypoints_3d'], out['pred_cam'], out['pred_smpl_params'], ...

Python here is host glue only (tensor allocation, dict packing, H2D of the input): every operation between
batch['img'] and the output tensors runs in libtokenhmr_b200.so through the C ABI, and fails loudly
if the library is missing.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, lib
from .config import TokenHMRConfig
from .ops import SMPLModel
from .weights import PackedWeights, make_config_struct


class _SmplFacade:
    """`model.smpl.faces` is read by demo.py:52."""

    def __init__(self, model: SMPLModel):
        self._model = model
        self.faces = model.faces


class TokenHMREngine(nn.Module):
    def __init__(self, cfg: TokenHMRConfig, state_dict: Dict[str, torch.Tensor], smpl: Dict[str, torch.Tensor],
                 device: str | torch.device = "cuda:0", max_batch: int = 64, use_cuda_graph: bool = True):
        super().__init__()
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.ThmrError("TokenHMREngine needs a CUDA device (there is no CPU fallback)")
        lib()  # fail now if the shared library is missing
        with torch.cuda.device(self.device):
            self.weights = PackedWeights(state_dict, cfg, self.device)
            self.smpl_model = SMPLModel(smpl, self.device)
            self.smpl = _SmplFacade(self.smpl_model)
            self._cfg_struct = make_config_struct(cfg)
            h = ctypes.c_void_p()
            check(lib().thmr_engine_create(ctypes.byref(self._cfg_struct), ctypes.byref(self.weights.struct),
                                           self.smpl_model.handle, ctypes.byref(h)))
            self._h = h
        self.max_batch = max_batch
        self.use_cuda_graph = use_cuda_graph
        self._bufs: Dict[int, dict] = {}     # per batch size: static buffers, workspace, optional graph
        self._dummy = nn.Parameter(torch.zeros(1), requires_grad=False)  # so .to()/.eval() behave like a Module

    # ------------------------------------------------------------------------------------------
    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().thmr_engine_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def to(self, *args, **kwargs):  # weights are already resident on self.device
        return self

    def num_launches(self) -> int:
        return lib().thmr_engine_num_launches(self._h)

    def _state(self, B: int, taps: bool, slot: int = 0) -> dict:
        key = (B, int(taps), slot)
        st = self._bufs.get(key)
        if st is not None:
            return st
        c, dev = self.cfg, self.device
        nj = 25 + self.smpl_model.n_extra
        f = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        t = {
            "img": f(B, 3, c.image_size, c.image_size),
            "cls_logits_softmax": f(B, c.token_num, c.token_class_num),
            "pred_cam": f(B, 3), "rotmats": f(B, c.num_joints, 3, 3), "betas": f(B, c.num_betas),
            "pred_cam_t": f(B, 3), "focal_length": f(B, 2), "pred_keypoints_3d": f(B, nj, 3),
            "pred_vertices": f(B, self.smpl_model.num_verts, 3), "pred_keypoints_2d": f(B, nj, 2),
        }
        if taps:
            t.update({"vit_tokens": f(B, c.num_tokens, c.vit_dim), "token_out": f(B, c.dec_dim), "pose6d": f(B, 144)})
        outs = _lib.Outputs()
        for name, _ in _lib.Outputs._fields_:
            if name in t:
                setattr(outs, name, t[name].data_ptr())
        nbytes = lib().thmr_engine_workspace_bytes(self._h, B)
        ws = torch.empty(nbytes + 1024, device=dev, dtype=torch.uint8)
        off = (-ws.data_ptr()) % 1024
        st = {"t": t, "outs": outs, "ws": ws, "ws_ptr": ws.data_ptr() + off, "graph": None, "warm": False}
        self._bufs[key] = st
        return st

    def _launch(self, st: dict, B: int) -> None:
        check(lib().thmr_engine_forward(self._h, st["t"]["img"].data_ptr(), B, ctypes.byref(st["outs"]), st["ws_ptr"],
                                        torch.cuda.current_stream().cuda_stream))

    @torch.no_grad()
    def forward(self, batch: Dict, return_taps: bool = False, slot: int = 0) -> Dict:
        """TokenHMR.forward: only batch['img'] is read (tokenhmr.py:146).  `slot` selects an independent set of
        input / output / workspace buffers (and CUDA graph), so that a caller can have several forwards in flight
        (TokenHMRPipeline); the returned tensors alias that slot's buffers until its next forward."""
        img = batch["img"]
        if img.dim() != 4 or img.shape[1] != 3 or img.shape[2] != self.cfg.image_size or img.shape[3] != self.cfg.image_size:
            raise _lib.ThmrError(f"batch['img'] must be (B,3,{self.cfg.image_size},{self.cfg.image_size}), got {tuple(img.shape)}")
        B = img.shape[0]
        with torch.cuda.device(self.device):
            st = self._state(B, return_taps, slot)
            st["t"]["img"].copy_(img.to(torch.float32), non_blocking=True)     # H2D (or D2D) of the batch
            if self.use_cuda_graph:
                if st["graph"] is None:
                    if not st["warm"]:
                        self._launch(st, B)       # eager call builds the plans and configures the kernels
                        torch.cuda.current_stream().synchronize()
                        st["warm"] = True
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._launch(st, B)
                    st["graph"] = g
                st["graph"].replay()
            else:
                self._launch(st, B)
            t = st["t"]
            rot = t["rotmats"]
            out = {
                "cls_logits_softmax": t["cls_logits_softmax"],
                "pred_cam": t["pred_cam"],
                "pred_smpl_params": {"global_orient": rot[:, :1].clone(), "body_pose": rot[:, 1:].clone(),
                                     "betas": t["betas"].clone()},
                "pred_cam_t": t["pred_cam_t"],
                "focal_length": t["focal_length"],
                "pred_keypoints_3d": t["pred_keypoints_3d"],
                "pred_vertices": t["pred_vertices"],
                "pred_keypoints_2d": t["pred_keypoints_2d"],
            }
            if return_taps:
                out["_vit_tokens"], out["_token_out"], out["_pred_body_pose_6d"] = t["vit_tokens"], t["token_out"], t["pose6d"]
        return out

    @torch.no_grad()
    def profile(self, img: torch.Tensor) -> list:
        """One eager forward with a CUDA event between launch groups.  Returns [(label, ms, flops, bytes), ...]
        (labels and algorithmic work come from the engine: thmr_engine_step_info)."""
        B = img.shape[0]
        with torch.cuda.device(self.device):
            st = self._state(B, False)
            st["t"]["img"].copy_(img.to(torch.float32))
            if not st["warm"]:
                self._launch(st, B)
                torch.cuda.current_stream().synchronize()
                st["warm"] = True
            n = lib().thmr_engine_num_steps(self._h)
            ms = (ctypes.c_float * n)()
            check(lib().thmr_engine_profile(self._h, st["t"]["img"].data_ptr(), B, ctypes.byref(st["outs"]),
                                            st["ws_ptr"], torch.cuda.current_stream().cuda_stream, ms, n))
            out = []
            for i in range(n):
                name, fl, by = ctypes.c_char_p(), ctypes.c_double(), ctypes.c_double()
                check(lib().thmr_engine_step_info(self._h, i, ctypes.byref(name), ctypes.byref(fl), ctypes.byref(by)))
                out.append((name.value.decode(), float(ms[i]), fl.value, by.value))
            return out

    @torch.no_grad()
    def backbone(self, img: torch.Tensor) -> torch.Tensor:
        """ViT.forward (vit.py:341-343): (B,3,256,256) -> (B,1280,16,12) like the reference backbone."""
        B = img.shape[0]
        with torch.cuda.device(self.device):
            st = self._state(B, True)
            st["t"]["img"].copy_(img.to(torch.float32), non_blocking=True)
            check(lib().thmr_engine_vit_forward(self._h, st["t"]["img"].data_ptr(), B, st["t"]["vit_tokens"].data_ptr(),
                                                st["ws_ptr"], torch.cuda.current_stream().cuda_stream))
            tok = st["t"]["vit_tokens"]
            gh, gw = self.cfg.grid_h, self.cfg.grid_w
            return tok.permute(0, 2, 1).reshape(B, self.cfg.vit_dim, gh, gw).contiguous()


class TokenHMRPipeline:
    """Double-buffered host -> device -> host driver for streams of batches (eval.py / track.py style loops):

        pipe = TokenHMRPipeline(model, read_back=("pred_vertices", "pred_keypoints_3d", "pred_cam", "pred_cam_t"))
        t0 = pipe.submit(batch0)            # H2D of the pinned host batch on the copy stream, forward, D2H: all async
        t1 = pipe.submit(batch1)            # its H2D overlaps the forward of batch0
        out0 = pipe.result(t0)              # pinned host tensors (valid until the slot is reused, `depth` submits later)

    Nothing is skipped: every submit copies its own input and every result is read back; only the waiting is moved, so
    the copy engines work while the SMs run the previous batch.  `post(out)` (optional) runs on the compute stream between
    the forward and the read-back (e.g. the all-gather of a sharded model)."""

    def __init__(self, model: "TokenHMREngine", depth: int = 2, read_back=("pred_vertices", "pred_keypoints_3d",
                                                                         "pred_cam", "pred_cam_t"), post=None):
        self.model, self.depth, self.read_back, self.post = model, int(depth), tuple(read_back), post
        with torch.cuda.device(model.device):
            self.copy_stream = torch.cuda.Stream(model.device)
            self.compute_stream = torch.cuda.Stream(model.device)
            self._copied = [torch.cuda.Event() for _ in range(self.depth)]
            self._done = [torch.cuda.Event() for _ in range(self.depth)]
        self._used = [False] * self.depth
        self._host = [dict() for _ in range(self.depth)]
        self._n = 0

    @torch.no_grad()
    def submit(self, batch: Dict) -> int:
        ticket = self._n
        slot = ticket % self.depth
        self._n += 1
        m = self.model
        img = batch["img"]
        with torch.cuda.device(m.device):
            st = m._state(img.shape[0], False, slot)
            with torch.cuda.stream(self.copy_stream):
                if self._used[slot]:
                    self.copy_stream.wait_event(self._done[slot])       # the slot's previous forward has consumed its input
                st["t"]["img"].copy_(img, non_blocking=True)
                self._copied[slot].record(self.copy_stream)
            with torch.cuda.stream(self.compute_stream):
                self.compute_stream.wait_event(self._copied[slot])
                out = m.forward({"img": st["t"]["img"]}, slot=slot)   # the D2D self-copy of the input is a no-op
                if self.post is not None:
                    out = self.post(out)
                host = self._host[slot]
                for k in self.read_back:
                    if k not in host or host[k].shape != out[k].shape:
                        host[k] = torch.empty(out[k].shape, dtype=out[k].dtype).pin_memory()
                    host[k].copy_(out[k], non_blocking=True)
                self._done[slot].record(self.compute_stream)
            self._used[slot] = True
        return ticket

    def result(self, ticket: int) -> Dict[str, torch.Tensor]:
        slot = ticket % self.depth
        self._done[slot].synchronize()
        return self._host[slot]


def load_tokenhmr(state_dict: Dict[str, torch.Tensor], smpl: Dict[str, torch.Tensor],
                  cfg: Optional[TokenHMRConfig] = None, device: str = "cuda:0", **kw):
    """Counterpart of lib.models.load_tokenhmr (tokenhmr/lib/models/__init__.py:3-26): returns (model, cfg).
    `state_dict` uses the reference key names (see tokenhmr_b200.weights.strip_checkpoint)."""
    cfg = cfg or TokenHMRConfig()
    model = TokenHMREngine(cfg, state_dict, smpl, device=device, **kw)
    return model, cfg
