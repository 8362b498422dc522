"""Direct-definition NumPy convolutions for the golden-vector generator -- deliberately INDEPENDENT of oracle/nets.py
(which uses torch conv3d / conv_transpose3d + pad / crop rules): here every output is written from TensorFlow's documented
definitions, tap by tap, so the goldens pin the SAME-padding and Conv3DTranspose arithmetic as well as the layer graph.

tf.nn.convolution, padding='SAME', stride s (tensorflow/python/ops/nn_ops.py docstring, "SAME" rule of
tensorflow/core/framework/common_shape_fns.cc GetWindowedOutputSizeVerbose):
    out_size  = ceil(n / s)
    pad_total = max((out_size - 1) * s + k - n, 0);  pad_before = pad_total // 2
    out[o] = sum_t  in[o*s + t - pad_before] * W[t]          (zero outside the input; W is NOT flipped)
tf.nn.conv3d_transpose == the gradient of that convolution w.r.t. its input (nn_ops.py: conv3d_backprop_input), for the
forward convolution that maps the OUTPUT shape (n*s) to the input shape (n):
    out[i*s + t - pad_before(n*s, k, s)] += x[i] . W[t]       (Keras kernel layout [kd,kh,kw,Cout,Cin], contraction over Cin)
"""
import math

import numpy as np


def same_pad_before(n, k, s):
    out = math.ceil(n / s)
    return max((out - 1) * s + k - n, 0) // 2


def conv3d_same(x, kernel, bias=None, stride=1):
    """x [B,D,H,W,Cin], kernel [k,k,k,Cin,Cout] -> [B,ceil(D/s),..,Cout] in x.dtype."""
    B, D, H, Wd, Ci = x.shape
    k = kernel.shape[0]
    s = stride
    od, oh, ow = math.ceil(D / s), math.ceil(H / s), math.ceil(Wd / s)
    pb = [same_pad_before(n, k, s) for n in (D, H, Wd)]
    out = np.zeros((B, od, oh, ow, kernel.shape[4]), x.dtype)
    kernel = kernel.astype(x.dtype)
    for tz in range(k):
        for ty in range(k):
            for tx in range(k):
                # output positions o with 0 <= o*s + t - pb < n
                rng = []
                for t, p, n, on in ((tz, pb[0], D, od), (ty, pb[1], H, oh), (tx, pb[2], Wd, ow)):
                    lo = max(0, math.ceil((p - t) / s))
                    hi = min(on - 1, (n - 1 + p - t) // s)
                    rng.append((lo, hi, t - p))
                if any(lo > hi for lo, hi, _ in rng):
                    continue
                (z0, z1, dz), (y0, y1, dy), (x0, x1, dx) = rng
                src = x[:, z0 * s + dz:z1 * s + dz + 1:s, y0 * s + dy:y1 * s + dy + 1:s, x0 * s + dx:x1 * s + dx + 1:s, :]
                out[:, z0:z1 + 1, y0:y1 + 1, x0:x1 + 1, :] += src @ kernel[tz, ty, tx]
    if bias is not None:
        out += bias.astype(x.dtype)
    return out


def conv3d_transpose_same(x, kernel, bias=None, stride=2):
    """x [B,D,H,W,Cin], kernel [k,k,k,Cout,Cin] -> [B,D*s,H*s,W*s,Cout] in x.dtype."""
    B, D, H, Wd, Ci = x.shape
    k = kernel.shape[0]
    s = stride
    Co = kernel.shape[3]
    on = (D * s, H * s, Wd * s)
    pb = [same_pad_before(n, k, s) for n in on]
    out = np.zeros((B,) + on + (Co,), x.dtype)
    kernel = kernel.astype(x.dtype)
    for tz in range(k):
        for ty in range(k):
            for tx in range(k):
                rng = []
                for t, p, n, o in ((tz, pb[0], D, on[0]), (ty, pb[1], H, on[1]), (tx, pb[2], Wd, on[2])):
                    # input positions i with 0 <= i*s + t - p < o
                    lo = max(0, math.ceil((p - t) / s))
                    hi = min(n - 1, (o - 1 + p - t) // s)
                    rng.append((lo, hi, t - p))
                if any(lo > hi for lo, hi, _ in rng):
                    continue
                (z0, z1, dz), (y0, y1, dy), (x0, x1, dx) = rng
                contrib = x[:, z0:z1 + 1, y0:y1 + 1, x0:x1 + 1, :] @ kernel[tz, ty, tx].T          # [.., Cin] @ [Cin, Cout]
                out[:, z0 * s + dz:z1 * s + dz + 1:s, y0 * s + dy:y1 * s + dy + 1:s, x0 * s + dx:x1 * s + dx + 1:s, :] += contrib
    if bias is not None:
        out += bias.astype(x.dtype)
    return out
