"""ORACLE (test infrastructure, not product code): CPU fp32 restatement of the reference
``BGModel`` + ``hardnet`` forward, written functionally over a plain state_dict.

Follows
  /root/reference/panoptic_forecasting/models/bg/bg_model.py:50-71,91-102  (one-hot, depth norm, predict)
  /root/reference/panoptic_forecasting/models/bg/hardnet.py:16-25          (ConvLayer = conv+BN+ReLU)
  .../hardnet.py:177-194,220-240                                            (HarDBlock links / forward)
  .../hardnet.py:243-258                                                    (TransitionUp)
  .../hardnet.py:262-327,353-387                                            (topology, forward)
It is a floating-point path, so the restatement uses the same torch fp32 CPU operators the
reference calls (F.conv2d, F.batch_norm, F.avg_pool2d, F.interpolate); tolerance vs the CUDA
path is stated in the tests (<= 1e-3 relative on logits, north_star).

Pinning: compared against the unmodified reference BGModel run in the build container with
the same seeded state_dict (tests/golden/make_golden.py -> tests/golden/bg_*.npz,
tests/test_oracle.py).  The reference ships no tests/golden vectors of its own.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
import torch
import torch.nn.functional as F

FIRST_CH = [16, 24, 32, 48]
CH_LIST = [64, 96, 160, 224, 320]
GRMUL = 1.7
GR = [10, 16, 18, 24, 32]
N_LAYERS = [4, 4, 8, 8, 8]


def get_link(layer, base_ch, growth_rate, grmul):
    """hardnet.py:177-194."""
    if layer == 0:
        return base_ch, 0, []
    out_channels = growth_rate
    link = []
    for i in range(10):
        dv = 2 ** i
        if layer % dv == 0:
            link.append(layer - dv)
            if i > 0:
                out_channels *= grmul
    out_channels = int(int(out_channels + 1) / 2) * 2
    in_channels = 0
    for i in link:
        in_channels += get_link(i, base_ch, growth_rate, grmul)[0]
    return out_channels, in_channels, link


def state_dict_shapes(num_classes=11, num_inputs=3, use_depth=True):
    """Names and shapes of the reference BGModel's state_dict (418 entries for the bg configs), derived from the
    topology constants above (hardnet.py:262-339, bg_model.py:17-48) -- so that callers that must not import the
    product package (bench.py --impl reference on the port) can build seeded weights."""
    shapes = {}
    if use_depth:
        shapes["depth_mean"] = (1,)
        shapes["depth_std"] = (1,)

    def conv(prefix, cin, cout, k):
        shapes[prefix + ".conv.weight"] = (cout, cin, k, k)
        for n in ("weight", "bias", "running_mean", "running_var"):
            shapes[prefix + ".norm." + n] = (cout,)
        shapes[prefix + ".norm.num_batches_tracked"] = ()

    def block(prefix, in_ch, gr, n_layers):
        out_ch = 0
        for l in range(n_layers):
            outch, inch, _ = get_link(l + 1, in_ch, gr, GRMUL)
            conv("%s.layers.%d" % (prefix, l), inch, outch, 3)
            if l % 2 == 0 or l == n_layers - 1:
                out_ch += outch
        return out_ch

    cin0 = (num_classes + (1 if use_depth else 0)) * num_inputs
    p = "model."
    conv(p + "base.0", cin0, FIRST_CH[0], 3)
    conv(p + "base.1", FIRST_CH[0], FIRST_CH[1], 3)
    conv(p + "base.2", FIRST_CH[1], FIRST_CH[2], 3)
    conv(p + "base.3", FIRST_CH[2], FIRST_CH[3], 3)
    idx, ch, skips = 4, FIRST_CH[3], []
    blks = len(N_LAYERS)
    for i in range(blks):
        ch = block(p + "base.%d" % idx, ch, GR[i], N_LAYERS[i])
        idx += 1
        if i < blks - 1:
            skips.append(ch)
        conv(p + "base.%d" % idx, ch, CH_LIST[i], 1)
        idx += 1
        ch = CH_LIST[i]
        if i < blks - 1:
            idx += 1                                   # AvgPool2d holds no parameters
    for j in range(blks - 1):
        i = blks - 2 - j
        cat = ch + skips.pop()
        conv(p + "conv1x1_up.%d" % j, cat, cat // 2, 1)
        ch = block(p + "denseBlocksUp.%d" % j, cat // 2, GR[i], N_LAYERS[i])
    shapes[p + "finalConv.weight"] = (num_classes, ch, 1, 1)
    shapes[p + "finalConv.bias"] = (num_classes,)
    return shapes


def conv_layer(sd, prefix, x, kernel, stride=1):
    """hardnet.py:16-25 in eval mode."""
    x = F.conv2d(x, sd[prefix + ".conv.weight"], None, stride, kernel // 2)
    x = F.batch_norm(x, sd[prefix + ".norm.running_mean"], sd[prefix + ".norm.running_var"],
                     sd[prefix + ".norm.weight"], sd[prefix + ".norm.bias"], False, 0.0, 1e-5)
    return F.relu(x)


def hard_block(sd, prefix, x, in_ch, gr, n_layers):
    """hardnet.py:220-240 (keepBase=False)."""
    layers = [x]
    for l in range(n_layers):
        _, _, link = get_link(l + 1, in_ch, gr, GRMUL)
        tin = [layers[i] for i in link]
        xin = torch.cat(tin, 1) if len(tin) > 1 else tin[0]
        layers.append(conv_layer(sd, "%s.layers.%d" % (prefix, l), xin, 3))
    t = len(layers)
    outs = [layers[i] for i in range(t) if i == t - 1 or i % 2 == 1]
    return torch.cat(outs, 1)


def hardnet_forward(sd, x, final_size=None, prefix="model."):
    """hardnet.py:353-387. Returns (final_out, quarter_res_out)."""
    size_in = x.shape
    p = prefix
    x = conv_layer(sd, p + "base.0", x, 3, 2)
    x = conv_layer(sd, p + "base.1", x, 3)
    x = conv_layer(sd, p + "base.2", x, 3, 2)
    x = conv_layer(sd, p + "base.3", x, 3)
    idx = 4
    ch = FIRST_CH[3]
    skips = []
    blks = len(N_LAYERS)
    for i in range(blks):
        x = hard_block(sd, p + "base.%d" % idx, x, ch, GR[i], N_LAYERS[i])
        idx += 1
        if i < blks - 1:
            skips.append(x)
        x = conv_layer(sd, p + "base.%d" % idx, x, 1)
        idx += 1
        ch = CH_LIST[i]
        if i < blks - 1:
            x = F.avg_pool2d(x, 2, 2)
            idx += 1
    out = x
    n_blocks = blks - 1
    for j in range(n_blocks):
        i = n_blocks - 1 - j
        skip = skips.pop()
        out = F.interpolate(out, size=(skip.size(2), skip.size(3)), mode="bilinear", align_corners=True)
        out = torch.cat([out, skip], 1)
        out = conv_layer(sd, p + "conv1x1_up.%d" % j, out, 1)
        out = hard_block(sd, p + "denseBlocksUp.%d" % j, out, out.shape[1], GR[i], N_LAYERS[i])
    out = F.conv2d(out, sd[p + "finalConv.weight"], sd[p + "finalConv.bias"])
    size = final_size if final_size is not None else (size_in[2], size_in[3])
    final_out = F.interpolate(out, size=size, mode="bilinear", align_corners=True)
    return final_out, out


def bg_inputs_to_planes(sd, seg, depth, depth_mask, num_classes=11):
    """bg_model.py:50-59,61-69: [b,t,H,W] int labels -> [b, t*11 + t, H, W] float planes
    (channel = frame*11 + class; then the t normalised masked depth planes)."""
    seg = seg.long()
    m = seg < num_classes
    seg = torch.where(m, seg, torch.zeros_like(seg))
    oh = F.one_hot(seg, num_classes) * m.unsqueeze(-1)
    oh = oh.permute(0, 1, 4, 2, 3).float()
    b, t, c, h, w = oh.shape
    x = oh.reshape(b, t * c, h, w)
    dn = (depth - sd["depth_mean"]) / sd["depth_std"]
    dn = dn * depth_mask
    return torch.cat([x, dn], 1)


def predict(sd, inputs, final_size=None):
    """bg_model.py:91-102."""
    with torch.no_grad():
        x = bg_inputs_to_planes(sd, inputs["seg"], inputs["depth"], inputs["depth_mask"])
        logits, quarter = hardnet_forward(sd, x, final_size)
        return {"seg": logits.argmax(1), "logits": logits, "orig_size_logits": quarter}


def predict_dense(sd, inputs, final_size=None):
    """bg_model.py:61-69,91-102 with `convert2onehot` off: inputs['seg'] already is a float [b,t,C,H,W] tensor of
    per-class planes; it is flattened to [b, t*C, H, W] and the t normalised masked depth planes are appended."""
    with torch.no_grad():
        inps = inputs["seg"].float()
        b, t, c, h, w = inps.shape
        x = inps.reshape(b, t * c, h, w)
        dn = (inputs["depth"] - sd["depth_mean"]) / sd["depth_std"]
        dn = dn * inputs["depth_mask"]
        logits, quarter = hardnet_forward(sd, torch.cat([x, dn], 1), final_size)
        return {"seg": logits.argmax(1), "logits": logits, "orig_size_logits": quarter}


def loss(sd, inputs, labels, final_size=None, dense=False):
    """bg_model.py:73-89: cross entropy (ignore_index 255) and pixel accuracy of the full-size logits."""
    out = (predict_dense if dense else predict)(sd, inputs, final_size)
    target = labels["seg"].long()
    ls = F.cross_entropy(out["logits"], target, ignore_index=255)
    correct = (out["logits"].argmax(1) == target).sum()
    total = (target != 255).sum()
    return {"loss": ls, "accuracy": correct.float() / total.float()}
