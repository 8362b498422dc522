"""One-off calibration of pcgcv1_b200.weights.CALIBRATION (run in the build container).

Runs the ORACLE nets (CPU) on a few synthetic vox10 cubes and prints final-layer gains such
that |y|max ~ 12, |z|max ~ 10, scale in ~[0.05, 3], loc spread similar to y, logits std ~ 3.
Dev tool: not imported by the product, tests or bench.
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pcgcv1_b200 import weights as W, synthetic as S
from oracle import nets

def stats(name, a):
    a = np.asarray(a)
    print("  %-8s min %9.3f max %9.3f mean %8.3f std %8.3f" % (name, a.min(), a.max(), a.mean(), a.std()))

def run(model, cal):
    w = W.synthetic_weights(model, calibration=cal)
    cubes, pos, nums = S.workload("vox10")
    sel = np.argsort(nums)[[0, len(nums)//4, len(nums)//2, 3*len(nums)//4, len(nums)-1]]
    x = cubes[sel].astype(np.float32)
    y = nets.run_net(model, "analysis", x, W.net_weights(w, "analysis_transform"))
    stats("y", y)
    out = {"y": y}
    if model == "voxception":
        z = nets.run_net(model, "hyper_encoder", y, W.net_weights(w, "hyper_encoder"))
        stats("z", z)
        loc, scale = nets.run_net(model, "hyper_decoder", np.rint(z), W.net_weights(w, "hyper_decoder"))
        stats("loc", loc); stats("scale", scale)
        out.update(z=z, loc=loc, scale=scale)
    xr = nets.run_net(model, "synthesis", np.rint(y), W.net_weights(w, "synthesis_transform"))
    stats("logits", xr)
    out["logits"] = xr
    return out

if __name__ == "__main__":
    model = sys.argv[1] if len(sys.argv) > 1 else "voxception"
    cal = {k: (1.0, 0.0) for k in W.CALIBRATION[model]}
    print("uncalibrated:")
    o = run(model, cal)
    tgt = {"y": 12.0, "z": 10.0, "loc": 8.0, "scale": 2.5, "logits": 12.0}
    key = {"y": "analysis_transform/conv_out" if model == "voxception" else "analysis_transform/conv_3",
           "z": "hyper_encoder/conv3", "loc": "hyper_decoder/deconv4_1", "scale": "hyper_decoder/deconv4_2",
           "logits": "synthesis_transform/deconv_out" if model == "voxception" else "synthesis_transform/deconv_3"}
    # two passes: downstream nets see the calibrated upstream latents
    for it in range(3):
        for name in ("y", "z", "loc", "scale", "logits"):
            if name not in o: continue
            g = tgt[name] / np.abs(o[name]).max()
            k = key[name]
            cal[k] = (cal[k][0] * g, cal[k][1])
        print("pass", it, cal)
        o = run(model, cal)
    print("CALIBRATION[%r] =" % model, {k: (round(v[0], 5), v[1]) for k, v in cal.items()})
