"""Extracts the BVH cross-check dumps of SURVEY.md Appendix B into JSON fixtures.

Those dumps were produced during the survey by an independent float32 restatement of
scene.rs / mesh.rs / bvh.rs (not by the reference binary, which cannot run here), so they
are a second opinion, not reference output. Run from the repo root:
    python tests/golden/make_bvh_golden.py
"""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def section(text, title):
    start = text.index(title)
    block = text[start:text.index("```", text.index("```", start) + 3)]
    return block


def parse(block):
    nodes_txt = block[block.index("nodes:"):block.index("primitives")]
    nodes = [[int(i), kind, int(a), int(b)] for i, kind, a, b in
             re.findall(r"(\d+) ([NL])\((\d+),(\d+)\)", nodes_txt)]
    prim_txt = block[block.index("primitives"):]
    prim_txt = prim_txt[:prim_txt.index("root bounds")]
    prims = [[int(a), int(b)] for a, b in re.findall(r"\((\d+),(\d+)\)", prim_txt)]
    rb = re.search(r"root bounds: min \(([^)]*)\)\s+max \(([^)]*)\)", block)
    root = [[float(x) for x in rb.group(1).split(",")], [float(x) for x in rb.group(2).split(",")]]
    return {"nodes": nodes, "primitives": prims, "root_bounds": root,
            "note": "N(second_child, split_axis) interior / L(first_slot, count) leaf"}


def main():
    text = open(os.path.join(ROOT, "SURVEY.md")).read()
    text = text[text.index("## Appendix B"):]
    for name, title in (("house", "**`assets/scenes/house.toml`**"),
                        ("default", "**`assets/scenes/default.toml`**")):
        data = parse(section(text, title))
        with open(os.path.join(ROOT, "tests", "golden", f"bvh_{name}.json"), "w") as fh:
            json.dump(data, fh, separators=(",", ":"))
        print(name, len(data["nodes"]), "nodes", len(data["primitives"]), "primitives")


if __name__ == "__main__":
    main()
