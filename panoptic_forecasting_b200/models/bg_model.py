"""BGModel on the B200 path: same ctor / forward / predict / load surface and the same 418
state_dict keys as the reference (panoptic_forecasting/models/bg/bg_model.py:17-102,
panoptic_forecasting/models/bg/hardnet.py:262-339), but forward() is one call into
libpf_b200.so (pf_bgnet_forward): labels are consumed as uint8 (no one-hot tensor), BatchNorm is
folded at load time, concatenations are channel slices of one NHWC arena, and the x4 bilinear
upsample + argmax is fused.  There is no torch/CPU fallback for the forward pass.
"""
import ctypes as C

import torch
from torch import nn

from .. import _lib
from .base_model import BaseModel


class _ConvBN(nn.Module):
    """Parameter holder with the reference ConvLayer's key names (hardnet.py:16-22)."""

    def __init__(self, cin, cout, k):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, kernel_size=k, padding=k // 2, bias=False)
        self.norm = nn.BatchNorm2d(cout)


def _insert(root, dotted, module):
    """Registers `module` under nested ModuleDicts so its state_dict prefix equals `dotted`."""
    parts = dotted.split('.')
    cur = root
    for p in parts[:-1]:
        if p not in cur:
            cur[p] = nn.ModuleDict()
        cur = cur[p]
    cur[parts[-1]] = module


class BGModel(BaseModel):

    def __init__(self, params):
        super().__init__()
        self.num_classes = num_classes = params['data']['num_classes']
        self.use_depth_inps = params['model'].get('use_depth_inps')
        self.num_inputs = params['model'].get('num_inputs', 1)
        self.min_depth = params['data'].get('min_depth')
        self.max_depth = params['data'].get('max_depth')
        # True: `inps` are class ids (uint8 / int64 [b,t,H,W]); False / None: float per-class planes [b,t,C,H,W]
        # (bg_model.py:61-65) -- the dense first-conv kernel (pf_bgnet_forward_dense)
        self.convert2onehot = params['model'].get('convert2onehot')
        final_w = params['model'].get('final_w')
        final_h = params['model'].get('final_h')
        self.final_size = (final_h, final_w) if final_w is not None and final_h is not None else None
        if self.use_depth_inps:
            depth_norm_params = params['data'].get('depth_norm_params')
            if depth_norm_params is None:
                mean, std = torch.zeros(1), torch.zeros(1)
            else:
                mean, std = depth_norm_params
            self.depth_mean = nn.Parameter(torch.as_tensor(mean, dtype=torch.float32).reshape(1), requires_grad=False)
            self.depth_std = nn.Parameter(torch.as_tensor(std, dtype=torch.float32).reshape(1), requires_grad=False)
        # B200-path options (absent keys keep the reference's outputs)
        b200 = params['model'].get('b200', {}) or {}
        self.precision = {'fp32': 0, 'tc': 1}[b200.get('precision', 'tc')]
        self.return_logits = b200.get('return_logits', True)
        self.seg_dtype = b200.get('seg_dtype', 'int64')

        self._lib = _lib.lib()
        self._net = C.c_void_p()
        _lib.check(self._lib.pf_bgnet_create(C.byref(self._net), self.num_classes, self.num_inputs,
                                             1 if self.use_depth_inps else 0, self.precision), "pf_bgnet_create")
        # parameter tree with the reference's key names, shapes taken from the native plan
        self.model = nn.ModuleDict()
        self._conv_names = []
        n = self._lib.pf_bgnet_num_convs(self._net)
        info = _lib.ConvInfo()
        for i in range(n):
            _lib.check(self._lib.pf_bgnet_conv_info(self._net, i, C.byref(info)), "pf_bgnet_conv_info")
            name = info.name.decode()
            assert name.startswith('model.')
            holder = _ConvBN(info.cin, info.cout, info.ksize)
            if i == 0:
                # reference expand_first_layer (hardnet.py:329-332): mean of a 3-channel init, repeated
                w3 = nn.Conv2d(3, info.cout, 3).weight.data
                holder.conv.weight.data = w3.mean(1, keepdim=True).expand(-1, info.cin, -1, -1).clone()
            _insert(self.model, name[len('model.'):], holder)
            self._conv_names.append(name)
        _lib.check(self._lib.pf_bgnet_conv_info(self._net, n, C.byref(info)), "pf_bgnet_conv_info")
        self.model['finalConv'] = nn.Conv2d(info.cin, info.cout, kernel_size=1, bias=True)
        nn.init.kaiming_normal_(self.model['finalConv'].weight)     # hardnet.py:334-339
        self._uploaded_device = None
        self._dirty = True
        self._ws = None

    def __del__(self):
        try:
            if getattr(self, '_net', None) and self._net.value:
                self._lib.pf_bgnet_destroy(self._net)
                self._net = C.c_void_p()
        except Exception:
            pass

    # -- weights --------------------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict=True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._dirty = True
        return out

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self._dirty = True
        return out

    def mark_weights_changed(self):
        self._dirty = True

    def _upload(self, device):
        """Folds BatchNorm and packs/uploads every conv for the native kernels (load-time prepack)."""
        sd = {k: v.detach().to('cpu', torch.float32).contiguous() for k, v in self.state_dict().items()}
        with torch.cuda.device(device):
            for i, name in enumerate(self._conv_names):
                w = sd[name + '.conv.weight']
                g, b_ = sd[name + '.norm.weight'], sd[name + '.norm.bias']
                m, v = sd[name + '.norm.running_mean'], sd[name + '.norm.running_var']
                _lib.check(self._lib.pf_bgnet_load_conv(self._net, i, w.data_ptr(), g.data_ptr(), b_.data_ptr(),
                                                        m.data_ptr(), v.data_ptr(), 1e-5), "pf_bgnet_load_conv(%s)" % name)
            fw, fb = sd['model.finalConv.weight'], sd['model.finalConv.bias']
            _lib.check(self._lib.pf_bgnet_load_final(self._net, fw.data_ptr(), fb.data_ptr()), "pf_bgnet_load_final")
            if self.use_depth_inps:
                _lib.check(self._lib.pf_bgnet_set_depth_norm(self._net, float(sd['depth_mean'][0]),
                                                             float(sd['depth_std'][0])), "pf_bgnet_set_depth_norm")
        self._uploaded_device = device
        self._dirty = False

    # -- forward --------------------------------------------------------------------------------
    def _run(self, inps, depths, depth_masks, want_full, want_quarter, want_seg):
        if not inps.is_cuda:
            raise _lib.PFError("BGModel.forward needs CUDA tensors (no CPU fallback)")
        dev = inps.device
        if self.training:
            raise NotImplementedError("the B200 path is inference-only (BatchNorm is folded); call .eval()")
        if self._dirty or self._uploaded_device != dev:
            self._upload(dev)
        dense = not self.convert2onehot
        if dense:
            if inps.dim() != 5 or inps.shape[2] != self.num_classes:
                raise ValueError("convert2onehot=False expects float planes [b, t, %d, H, W], got %s"
                                 % (self.num_classes, tuple(inps.shape)))
            b, t, _, H, W = inps.shape
        else:
            b, t, H, W = inps.shape
        if t != self.num_inputs:
            raise ValueError("expected %d input frames, got %d" % (self.num_inputs, t))
        if dense:
            labels = inps.to(torch.float32)
        elif inps.dtype == torch.uint8:
            labels = inps
        else:
            # ids outside [0, 255] (the reference's F.one_hot raises on negatives) take the all-zero one-hot row,
            # like every id >= num_classes (bg_model.py:54-56) -- never class 0
            labels = torch.where((inps < 0) | (inps > 255), torch.full_like(inps, 255), inps).to(torch.uint8)
        labels = labels.contiguous()
        depth_c = mask_c = None
        if self.use_depth_inps:
            depth_c = depths.to(torch.float32).contiguous()
            mask_c = depth_masks.contiguous()
            if mask_c.dtype.is_floating_point:
                # the reference multiplies by the mask (bg_model.py:68); the kernels take a 0/1 mask
                raise TypeError("depth_masks must be bool or an integer 0/1 mask, got %s" % mask_c.dtype)
            mask_c = mask_c.view(torch.uint8) if mask_c.dtype == torch.bool else mask_c.to(torch.uint8)
        fh, fw = self.final_size if self.final_size is not None else (H, W)
        nbytes = self._lib.pf_bgnet_workspace_bytes(self._net, b, H, W)
        if nbytes == 0:
            raise ValueError("unsupported input size %dx%d (H must be a multiple of 4, W of 16, both >= 64)" % (H, W))
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != dev:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        seg8 = seg64 = full = quarter = None
        if want_seg:
            if self.seg_dtype == 'uint8':
                seg8 = torch.empty((b, fh, fw), dtype=torch.uint8, device=dev)
            else:
                seg64 = torch.empty((b, fh, fw), dtype=torch.int64, device=dev)
        if want_full:
            full = torch.empty((b, self.num_classes, fh, fw), dtype=torch.float32, device=dev)
        if want_quarter:
            quarter = torch.empty((b, self.num_classes, H // 4, W // 4), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            fwd = self._lib.pf_bgnet_forward_dense if dense else self._lib.pf_bgnet_forward
            rc = fwd(self._net, labels.data_ptr(), _lib.ptr(depth_c), _lib.ptr(mask_c),
                     b, H, W, fh, fw, _lib.ptr(seg8), _lib.ptr(seg64), _lib.ptr(quarter),
                     _lib.ptr(full), self._ws.data_ptr(), self._ws.numel(), stream)
        _lib.check(rc, "pf_bgnet_forward_dense" if dense else "pf_bgnet_forward")
        return (seg8 if seg8 is not None else seg64), full, quarter

    def forward(self, inps, depths, depth_masks, return_orig_size=False):
        """Reference signature (bg_model.py:61-71): returns full-size logits (and quarter-res logits)."""
        _, full, quarter = self._run(inps, depths, depth_masks, True, return_orig_size, False)
        if return_orig_size:
            return full, quarter
        return full

    def loss(self, inputs, labels):
        """Reference signature (bg_model.py:73-89): cross entropy (ignore_index 255) and pixel accuracy of the
        full-size logits against labels['seg'].  Evaluation only: the logits come from the inference kernels
        (BatchNorm folded, no autograd graph), so the returned loss carries no gradient -- this is the number the
        reference's validation pass logs in .eval() mode; training stays with the reference (same state_dict keys)."""
        if self.training:
            raise NotImplementedError("the B200 path is inference-only (BatchNorm is folded): loss() gives the "
                                      "validation loss / accuracy in .eval() mode; train with the reference, then "
                                      ".load() the checkpoint here (same state_dict keys)")
        seg_labels = labels['seg']
        seg_preds = self(inputs['seg'], inputs.get('depth'), inputs.get('depth_mask'))
        seg_labels = seg_labels.to(seg_preds.device, torch.int64)
        seg_loss = nn.functional.cross_entropy(seg_preds, seg_labels, ignore_index=255)
        max_preds = seg_preds.argmax(1).long()
        correct = (max_preds == seg_labels).sum()
        total = (seg_labels != 255).sum()
        return {'loss': seg_loss, 'accuracy': correct.float() / total.float()}

    def predict(self, inputs, labels):
        """Reference signature (bg_model.py:91-102).  'logits' / 'orig_size_logits' are produced
        unless params['model']['b200']['return_logits'] is False (export path: label map only)."""
        inps = inputs['seg']
        depths = inputs.get('depth')
        depth_masks = inputs.get('depth_mask')
        seg, full, quarter = self._run(inps, depths, depth_masks, self.return_logits, self.return_logits, True)
        final_result = {'seg': seg}
        if self.return_logits:
            final_result['logits'] = full
            final_result['orig_size_logits'] = quarter
        return final_result
