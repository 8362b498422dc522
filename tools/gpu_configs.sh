#!/bin/bash
# other BASELINE configs: (2) factorized model_simple, (3) sparse vox12 sharded workload at N=1, (4) batch sweep
mkdir -p gpurun_out
timeout 900 python - <<'PY' 2>&1 | tee gpurun_out/configs.log
import time, json, numpy as np, torch
from pcgcv1_b200 import synthetic, transform, runtime
from pcgcv1_b200.dataprocess import inout_points
from pcgcv1_b200.models import model_voxception, model_simple
res = {}
# ---- config 2: factorized mode, model_simple, vox10
cubes, pos, nums = synthetic.workload("vox10")
pin = torch.from_numpy(cubes).pin_memory()
cs = runtime.get_codec("simple", "")
for it in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    s, mn, mx, shp = transform.compress_factorized(pin, model_simple, "")
    xs = transform.decompress_factorized(s.numpy(), mn.numpy(), mx.numpy(), shp.numpy(), model_simple, "")
    m = inout_points.select_voxels(xs, nums, 1.0, codec=cs, dtype="uint8")
    torch.cuda.synchronize(); dt = time.time() - t0
res["config2_factorized_simple_e2e_cubes_per_s"] = round(len(cubes) / dt, 1)
res["config2_bytes"] = len(s.numpy()); res["config2_range"] = [int(mn), int(mx)]
print("config 2:", res["config2_factorized_simple_e2e_cubes_per_s"], "cubes/s e2e; string", len(s.numpy()), "B; range", int(mn), int(mx))
# ---- config 3: vox12 sparse cloud (thousands of cubes), hyper mode, N=1
cubes, pos, nums = synthetic.workload("vox12")
pin = torch.from_numpy(cubes).pin_memory()
cv = runtime.get_codec("voxception", "")
for it in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    out = transform.compress_hyper(pin, model_voxception, "")
    host = [o.numpy() for o in out]
    t1 = time.time()
    xs = transform.decompress_hyper(*host, model_voxception, "")
    m = inout_points.select_voxels(xs, nums, 1.0, codec=cv, dtype="uint8")
    torch.cuda.synchronize(); dt = time.time() - t0
ymin, ymax = host[1].min(), host[2].max()
res["config3_vox12_cubes"] = len(cubes); res["config3_e2e_cubes_per_s"] = round(len(cubes) / dt, 1)
res["config3_y_range"] = [int(ymin), int(ymax)]; res["config3_bytes"] = int(sum(len(x) for x in host[0]) + len(host[4]))
print("config 3:", len(cubes), "cubes", res["config3_e2e_cubes_per_s"], "cubes/s e2e (compress %.2f s), y range" % (t1 - t0), ymin, ymax, "mask>=nums", bool((m.reshape(len(cubes), -1).sum(1) >= nums).all()))
del xs, m, out, host
torch.cuda.empty_cache()
# ---- config 4: batch sweep of analysis + synthesis
sw = {}
base, _ = synthetic.surface_cubes(8, seed=3)
for B in (8, 16, 32, 64, 128, 256, 512):
    x = cv.to_device(np.tile(base, (B // 8, 1, 1, 1, 1)))
    y = torch.randn(B, 16, 16, 16, 16, device=cv.dev) * 3
    for _ in range(5):
        cv.analysis(x); cv.synthesis(y)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    it = 20 if B <= 128 else 8
    a.record()
    for _ in range(it):
        cv.analysis(x); cv.synthesis(y)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / it
    sw[B] = {"ms": round(ms, 3), "cubes_per_s": round(B / ms * 1e3, 1), "tflops": round(B * 20.7996 / ms, 2)}
    print("config 4: B=%d %.2f ms -> %.0f cubes/s, %.1f TFLOP/s" % (B, ms, B / ms * 1e3, B * 20.7996 / ms))
res["config4_sweep"] = sw
json.dump(res, open("gpurun_out/configs.json", "w"), indent=1)
PY
