"""PCTransformModel on the B200 path: same ctor / predict surface as the reference
(panoptic_forecasting/models/pc_transform/pc_transform_model.py:18-150); the work is one call
into libpf_b200.so (pf_zsplat_forward)."""
import torch

from .. import _lib
from .base_model import BaseModel


class PCTransformModel(BaseModel):
    def __init__(self, params):
        super().__init__()
        self.ind = params['model'].get('only_this_ind')
        self.is_img = params['model'].get('is_img')
        self.debug = params['model'].get('debug')
        # extensions (absent keys keep reference behaviour)
        self.return_result2d = params['model'].get('return_result2d', True)
        self.label_lut = params['model'].get('label_lut')       # optional 256-entry remap
        self._ws = None
        self._lut_dev = None

    def _workspace(self, nbytes, device):
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != device:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return self._ws

    def predict(self, inputs, labels):
        K = inputs['intrinsics']
        extrinsics = inputs['extrinsics']
        depths = inputs['depth']
        depth_mask = inputs['depth_mask']
        target_T = inputs['target_T']
        segs = inputs['seg']
        if self.ind is not None:                      # reference :33-37
            depths = depths[:, self.ind:self.ind + 1]
            depth_mask = depth_mask[:, self.ind:self.ind + 1]
            target_T = target_T[:, self.ind:self.ind + 1]
            segs = segs[:, self.ind:self.ind + 1]
        if not depths.is_cuda:
            raise _lib.PFError("PCTransformModel.predict needs CUDA tensors (no CPU fallback)")
        dev = depths.device
        b, t, H, W = depths.shape
        payload = 3 if self.is_img else 1
        if segs.dtype != torch.uint8:
            raise TypeError("seg must be uint8 (reference dataset dtype), got %s" % segs.dtype)
        # The reference inverts K / E with torch.inverse (:51,:71).  Optional precomputed inverses
        # (e.g. from the CPU, for bit parity with a CPU run of the reference) may be supplied.
        K = K.to(dev, torch.float32).contiguous()
        E = extrinsics.to(dev, torch.float32).contiguous()
        Kinv = inputs['intrinsics_inv'].to(dev, torch.float32) if 'intrinsics_inv' in inputs else torch.inverse(K)
        Einv = inputs['extrinsics_inv'].to(dev, torch.float32) if 'extrinsics_inv' in inputs else torch.inverse(E)
        # torch.inverse returns column-major batches: materialise row-major copies and KEEP the
        # references alive until the launch is enqueued (a temporary would be recycled by the
        # caching allocator before the kernel reads it).
        Kinv = Kinv.contiguous()
        Einv = Einv.contiguous()
        depth_c = depths.to(torch.float32).contiguous()
        mask_c = depth_mask.to(torch.uint8).contiguous()
        seg_c = segs.contiguous()
        T_c = target_T.to(dev, torch.float32).contiguous()
        out_seg = torch.empty((b, H, W, 3) if self.is_img else (b, H, W), dtype=torch.uint8, device=dev)
        out_depth = torch.empty((b, H, W), dtype=torch.float32, device=dev)
        coords = torch.empty((b, t, H, W, 2), dtype=torch.int64, device=dev) if self.return_result2d else None
        L = _lib.lib()
        nbytes = L.pf_zsplat_workspace_bytes(b, t, H, W)
        ws = self._workspace(nbytes, dev)
        lut = None
        if self.label_lut is not None and not self.is_img:
            if self._lut_dev is None or self._lut_dev.device != dev:
                self._lut_dev = torch.as_tensor(self.label_lut, dtype=torch.uint8).to(dev).contiguous()
            lut = self._lut_dev
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = L.pf_zsplat_forward(depth_c.data_ptr(), mask_c.data_ptr(), seg_c.data_ptr(),
                                     K.data_ptr(), Kinv.data_ptr(), E.data_ptr(), Einv.data_ptr(), T_c.data_ptr(),
                                     b, t, H, W, payload, _lib.ptr(lut),
                                     out_seg.data_ptr(), out_depth.data_ptr(), _lib.ptr(coords),
                                     ws.data_ptr(), ws.numel(), stream)
        _lib.check(rc, "pf_zsplat_forward")
        result = {'seg': out_seg, 'depth': out_depth}
        if coords is not None:
            result['result2d'] = coords
        return result
