"""Drop-in for the reference's CLI ``test.py`` (:24-115): same positional arguments, flags, defaults, file naming
(``./compressed/<name>.*``, ``<name>_rec.ply``) and call sequence; the codec underneath is libpcgc_b200.so.

    python -m pcgcv1_b200.test compress  cloud.ply               --ckpt_dir=... [--mode=hyper|factorized]
    python -m pcgcv1_b200.test decompress compressed/cloud       --ckpt_dir=...

``--modelname`` takes the reference's module paths (``models.model_voxception``, ``models.model_simple``); ``--gpu`` is
accepted for compatibility -- there is no CPU path (``--gpu 0`` is an error rather than a silent slow run)."""
from __future__ import annotations

import argparse
import importlib
import os


def parse_args(argv=None):
    parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("command", choices=["compress", "decompress"],
                        help="What to do: 'compress' reads a point cloud (.ply format) and writes compressed binary files. "
                             "'decompress' reads binary files and reconstructs the point cloud (.ply format). "
                             "input and output filenames need to be provided for the latter. ")
    parser.add_argument("input", nargs="?", help="Input filename.")
    parser.add_argument("output", nargs="?", help="Output filename.")
    parser.add_argument("--mode", type=str, default='hyper', dest="mode", help='factorized entropy model or hyper prior')
    parser.add_argument("--modelname", default="models.model_voxception", dest="modelname", help="(model_simple, model_voxception)")
    parser.add_argument("--ckpt_dir", type=str, default='', dest="ckpt_dir", help='checkpoint')
    parser.add_argument("--scale", type=float, default=1.0, dest="scale", help="scaling factor.")
    parser.add_argument("--cube_size", type=int, default=64, dest="cube_size", help="size of partitioned cubes.")
    parser.add_argument("--min_num", type=int, default=64, dest="min_num", help="minimum number of points in a cube.")
    parser.add_argument("--rho", type=float, default=1.0, dest="rho", help="ratio of the numbers of output points to the number of input points.")
    parser.add_argument("--gpu", type=int, default=1, dest="gpu", help="use gpu (1) or not (0).")
    args = parser.parse_args(argv)
    print(args)
    return args


def main(argv=None):
    args = parse_args(argv)
    if args.gpu != 1:
        raise SystemExit("pcgcv1_b200 has no CPU path (--gpu 0): the codec is the CUDA library")
    if args.cube_size != 64:
        raise SystemExit("the transforms are built for 64^3 cubes (the reference's default and its checkpoints)")
    from .dataprocess.inout_bitstream import (read_binary_files_factorized, read_binary_files_hyper,
                                              write_binary_files_factorized, write_binary_files_hyper)
    from .process import postprocess, preprocess
    from .transform import compress_factorized, compress_hyper, decompress_factorized, decompress_hyper
    from . import runtime

    name = args.modelname if args.modelname.startswith("pcgcv1_b200.") else "pcgcv1_b200." + args.modelname
    model = importlib.import_module(name)
    codec = runtime.get_codec(model, args.ckpt_dir)

    if args.mode == "factorized":
        if args.command == "compress":
            cubes, cube_positions, points_numbers = preprocess(args.input, args.scale, args.cube_size, args.min_num, codec=codec)
            strings, min_v, max_v, shape = compress_factorized(cubes, model, args.ckpt_dir)
            if not args.output:
                args.output = os.path.split(args.input)[-1][:-4]
            write_binary_files_factorized(args.output, strings.numpy(), points_numbers, cube_positions, min_v.numpy(), max_v.numpy(),
                                          shape.numpy(), rootdir='./compressed')
        elif args.command == "decompress":
            rootdir, filename = os.path.split(args.input)
            if not args.output:
                args.output = filename + "_rec.ply"
            strings_d, points_numbers_d, cube_positions_d, min_v_d, max_v_d, shape_d = read_binary_files_factorized(filename, rootdir)
            cubes_d = decompress_factorized(strings_d, min_v_d, max_v_d, shape_d, model, args.ckpt_dir)
            postprocess(args.output, cubes_d, points_numbers_d, cube_positions_d, args.scale, args.cube_size, args.rho, codec=codec)

    if args.mode == "hyper":
        if args.command == "compress":
            if not args.output:
                args.output = os.path.split(args.input)[-1][:-4]
            cubes, cube_positions, points_numbers = preprocess(args.input, args.scale, args.cube_size, args.min_num, codec=codec)
            y_strings, y_min_vs, y_max_vs, y_shape, z_strings, z_min_v, z_max_v, z_shape = compress_hyper(cubes, model, args.ckpt_dir)
            write_binary_files_hyper(args.output, y_strings.numpy(), z_strings.numpy(), points_numbers, cube_positions,
                                     y_min_vs.numpy(), y_max_vs.numpy(), y_shape.numpy(), z_min_v.numpy(), z_max_v.numpy(),
                                     z_shape.numpy(), rootdir='./compressed')
        elif args.command == "decompress":
            rootdir, filename = os.path.split(args.input)
            if not args.output:
                args.output = filename + "_rec.ply"
            (y_strings_d, z_strings_d, points_numbers_d, cube_positions_d, y_min_vs_d, y_max_vs_d, y_shape_d, z_min_v_d, z_max_v_d,
             z_shape_d) = read_binary_files_hyper(filename, rootdir)
            cubes_d = decompress_hyper(y_strings_d, y_min_vs_d, y_max_vs_d, y_shape_d, z_strings_d, z_min_v_d, z_max_v_d, z_shape_d,
                                       model, args.ckpt_dir)
            postprocess(args.output, cubes_d, points_numbers_d, cube_positions_d, args.scale, args.cube_size, args.rho, codec=codec)


if __name__ == "__main__":
    main()
