#!/usr/bin/env python
"""rt_oracle — command line front of the CPU oracle (TEST INFRASTRUCTURE, see oracle.py).

Renders a scene with the CPU restatement of shader.wgsl and writes what the parity checks
compare: the BVH dump, the alias table, the first-hit slot map, the linear
accumulation buffer and the work counters. The product never imports this.

  python oracle/cli.py --scene assets/scenes/house.toml --width 320 --height 180 --spp 4 \\
      --bounces 8 --pfm out.pfm --first-hit hit.u32 --bvh bvh.bin --alias alias.bin --counters c.json
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(argv=None):
    ap = argparse.ArgumentParser(prog="rt_oracle", description=__doc__.split("\n\n")[0])
    ap.add_argument("--scene", required=True, help="scene TOML (assets/scenes/*.toml)")
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=360)
    ap.add_argument("--spp", type=int, default=1, help="samples 0 .. spp-1 (or --sample-begin ..)")
    ap.add_argument("--sample-begin", type=int, default=0)
    ap.add_argument("--bounces", type=int, default=10, help="MAX_BOUNCES (shader.wgsl:232)")
    ap.add_argument("--env", type=int, default=0, help="environment slot")
    ap.add_argument("--hdri", action="append", default=[], help=".hdr file(s); default: the synthetic sky")
    ap.add_argument("--prune", action="store_true", help="count work with closest-hit pruning (SURVEY §8d)")
    ap.add_argument("--threads", type=int, default=0, help="0 = all host cores")
    ap.add_argument("--pfm", help="linear accumulation buffer (RGB, float32)")
    ap.add_argument("--png", help="tone-mapped 8-bit image")
    ap.add_argument("--first-hit", help="u32 slot map of sample --sample-begin (row-major)")
    ap.add_argument("--bvh", help="BVH nodes (48-byte records) followed by primitive infos (8-byte records)")
    ap.add_argument("--alias", help="alias table of the environment slot (16-byte entries)")
    ap.add_argument("--counters", help="work counters as JSON")
    args = ap.parse_args(argv)

    import _pkg
    rt = _pkg.load()
    import oracle as orc

    scene = rt.Scene.load_toml(args.scene)
    envs = [rt.Environment.load_hdr(p) for p in args.hdri] or [rt.Environment.synthetic(2048, 1024)]
    env = rt.EnvironmentMaps(envs)
    o = orc.Oracle(scene, env, args.width, args.height, max_bounces=args.bounces,
                   environment_index=args.env, prune=args.prune)
    out = {"scene": args.scene, "width": args.width, "height": args.height, "spp": args.spp,
           "sample_begin": args.sample_begin, "bounces": args.bounces, "counts": scene.counts(),
           "bvh_depth": scene.bvh_depth}
    if args.bvh:
        with open(args.bvh, "wb") as fh:
            fh.write(np.ascontiguousarray(scene.bvh_nodes_words()).tobytes())
            fh.write(np.ascontiguousarray(scene.primitives_words()).tobytes())
    if args.alias:
        with open(args.alias, "wb") as fh:
            fh.write(np.ascontiguousarray(envs[args.env].alias).tobytes())
    if args.first_hit:
        slots, _t, c = o.first_hit(args.sample_begin, threads=args.threads)
        slots.tofile(args.first_hit)
        out["first_hit_counters"] = c
        out["first_hit_fraction"] = float((slots != 0xFFFFFFFF).mean())
    if args.pfm or args.png or args.counters:
        accum, c = o.render(args.sample_begin, args.spp, threads=args.threads)
        out["render_counters"] = c
        if args.pfm:
            rt.write_pfm(args.pfm, accum)
        if args.png:
            rt.write_png(args.png, orc.tonemap_rgb8(accum, args.sample_begin + args.spp))
    if args.counters:
        with open(args.counters, "w") as fh:
            json.dump(out, fh, indent=1)
    else:
        print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
