"""The rho search of the reference's evaluation driver: ``select_optimal_rho`` and
``cfg_post_process`` of ``eval_ablation_studies.py:152-205`` with their argument lists.  The rest of that script (the RD sweep
over checkpoints, its .ini bookkeeping of scales and checkpoint paths, the CSV / plots) is driver code around external files and
is not rebuilt.

What runs per candidate rho: ``process.postprocess`` -- top-k classification ON the GPU over the decoded logits, which stay
resident on the device across the whole search when ``cubes_d`` is the handle ``decompress_hyper`` returned (the reference re-sorts
every cube on the host for each rho), ordered extraction, .ply -- then ``myutils.pc_error_wrapper.pc_error`` (the MPEG binary when
present, the built-in D1 / D2 figures otherwise).

The search loop keeps the reference's exact control flow, including its quirk: the first candidate's PSNR never enters
``MAX_PSNR`` (it is initialised to 0 at i == 0), so the second candidate is always accepted and the search stops at the first
candidate from the third on whose PSNR drops below the best of candidates 2..i."""
from __future__ import annotations

from .myutils.pc_error_wrapper import pc_error
from .process import postprocess

RHOS_D1 = [0.8, 0.9, 1.0, 1.02, 1.05, 1.10, 1.15, 1.2, 1.25, 1.30, 1.40, 1.50, 1.75, 2.0, 2.5, 3.0]           # :180
RHOS_D2 = [1.0, 0.98, 0.95, 0.92, 0.90, 0.88, 0.85, 0.82, 0.80, 0.75, 0.70, 0.65, 0.50, 0.40, 0.30]           # :192
ITEM_D1, ITEM_D2 = "mseF,PSNR (p2point)", "mseF,PSNR (p2plane)"


def _scalar(results, item):
    v = results[item]
    return float(v.iloc[0]) if hasattr(v, "iloc") else float(v)


def select_optimal_rho(item, rhos, input_file, output_file, input_file_n, cubes_d, points_numbers_d, cube_positions_d, scale, cube_size, res,
                       post=postprocess, metric=pc_error):
    """eval_ablation_studies.py:152-174.  ``post`` / ``metric`` default to this package's ``postprocess`` / ``pc_error``."""
    optimal_rho = None
    for i, rho in enumerate(rhos):
        print("===== select rho =====")
        post(output_file, cubes_d, points_numbers_d, cube_positions_d, scale, cube_size, rho)
        results = metric(input_file, output_file, input_file_n, res, show=False)
        PSNR = _scalar(results, item)
        print("===== results: ", i, rho, item, PSNR)
        if i == 0:
            MAX_PSNR = 0
            optimal_rho = rho
        else:
            MAX_PSNR = max(PSNR, MAX_PSNR)
        if PSNR < MAX_PSNR:
            break
        else:
            optimal_rho = rho
    return optimal_rho


def cfg_post_process(config, config_file, rate, input_file, output_file, input_file_n, cubes_d, points_numbers_d, cube_positions_d, scale,
                     cube_size, res, post=postprocess, metric=pc_error):
    """eval_ablation_studies.py:177-203: rho_d1 / rho_d2 of section ``rate`` from the .ini when present, else searched and written back."""
    found = {}
    for key, item, rhos in (("rho_d1", ITEM_D1, RHOS_D1), ("rho_d2", ITEM_D2, RHOS_D2)):
        if config.has_option(rate, key):
            found[key] = float(config.get(rate, key))
        else:
            found[key] = select_optimal_rho(item, rhos, input_file, output_file, input_file_n, cubes_d, points_numbers_d, cube_positions_d,
                                            scale, cube_size, res, post=post, metric=metric)
            config.set(rate, key, str(found[key]))
            with open(config_file, "w") as f:
                config.write(f)
    return found["rho_d1"], found["rho_d2"]
