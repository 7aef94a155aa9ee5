"""C-ABI surface and environment inputs — CPU only (no compute calls on the backend)."""
import ctypes as C
import os
import re
import struct
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def _declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", text)))


def test_host_library_exports_every_declared_symbol(rt):
    lib = rt._ffi.host_lib()
    declared = _declared_functions("rt_host.h")
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"librt_host.so does not export {name}"
    assert set(declared) == set(rt._ffi.HOST_SIGNATURES), "ctypes table out of sync with rt_host.h"


def test_backend_library_loads_and_exports_every_declared_symbol(rt):
    # Loading must work without a GPU; only rt_create may fail there.
    lib = rt._ffi.backend_lib()
    declared = _declared_functions("rt_backend.h")
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"librt_backend.so does not export {name}"
    assert set(declared) == set(rt._ffi.BACKEND_SIGNATURES), "ctypes table out of sync with rt_backend.h"
    assert b"sm_100a" in lib.rt_backend_version()


def test_backend_has_no_cpu_path(rt, house, env_small, cuda_available):
    if cuda_available:
        pytest.skip("a GPU is present; the no-GPU failure mode is checked on CPU boxes")
    with pytest.raises(rt.RtError) as e:
        rt.State(house, env_small, 16, 16)
    assert e.value.status == rt._ffi.RT_ERR_CUDA
    assert "no CPU path" in str(e.value)


def test_backend_is_built_for_sm_100a_only():
    so = os.path.join(ROOT, "rsoderh-raytracing_b200", "librt_backend.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_product_never_references_the_oracle():
    # The oracle is test infrastructure: nothing under the package (or the Makefile's
    # product targets) may import, link or execute it.
    pkg = os.path.join(ROOT, "rsoderh-raytracing_b200")
    for base, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".hpp", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"
    for lib in ("librt_backend.so", "librt_host.so"):
        out = subprocess.run(["ldd", os.path.join(pkg, lib)], capture_output=True, text=True).stdout
        assert "oracle" not in out


def test_struct_layouts_match_wgsl_offsets(rt):
    f = rt._ffi
    assert (f.RtMaterial.roughness.offset, f.RtMaterial.metallic.offset, f.RtMaterial.emission.offset) == (12, 16, 32)
    assert (f.RtSphere.radius.offset, f.RtSphere.material_id.offset) == (12, 16)
    assert (f.RtPlane.normal.offset, f.RtPlane.base_change.offset, f.RtPlane.material_id.offset) == (16, 32, 80)
    assert (f.RtBvhNode.max.offset, f.RtBvhNode.primitives_or_second_child_index.offset,
            f.RtBvhNode.primitives_len.offset, f.RtBvhNode.split_axis.offset) == (16, 32, 36, 40)
    assert (f.RtCameraUniform.rot.offset, f.RtCameraUniform.fov_y.offset) == (16, 64)
    assert (f.RtAliasEntry.alias_index.offset, f.RtAliasEntry.pmf.offset) == (4, 8)


# ------------------------------------------------------------------ Radiance .hdr
def _rle_encode_scanline(rgbe_row):
    """New-style RLE (what Poly Haven files use), written independently of the decoder."""
    w = rgbe_row.shape[0]
    out = bytearray([2, 2, w >> 8, w & 255])
    for ch in range(4):
        data = rgbe_row[:, ch].tolist()
        i = 0
        while i < w:
            run = 1
            while i + run < w and run < 127 and data[i + run] == data[i]:
                run += 1
            if run >= 3:
                out += bytes([128 + run, data[i]])
                i += run
            else:
                j = i
                lit = []
                while j < w and len(lit) < 128:
                    r = 1
                    while j + r < w and r < 3 and data[j + r] == data[j]:
                        r += 1
                    if r >= 3:
                        break
                    lit.append(data[j])
                    j += 1
                out += bytes([len(lit)]) + bytes(lit)
                i = j
    return bytes(out)


def test_hdr_flat_and_rle_decode(rt, tmp_path):
    rng = np.random.default_rng(7)
    w, h = 48, 20
    rgbe = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    rgbe[:, 10:30] = rgbe[:, 10:11]          # long runs
    rgbe[3, :, 3] = 0                        # exponent 0 -> black
    expect = np.where(rgbe[..., 3:4] == 0, 0.0,
                      rgbe[..., :3].astype(np.float64) * np.exp2(rgbe[..., 3:4].astype(np.float64) - 136.0)
                      ).astype(np.float32)
    header = f"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y {h} +X {w}\n".encode()
    flat = tmp_path / "flat.hdr"
    flat.write_bytes(header + rgbe.tobytes())
    rle = tmp_path / "rle.hdr"
    rle.write_bytes(header + b"".join(_rle_encode_scanline(rgbe[y]) for y in range(h)))
    for p in (flat, rle):
        env = rt.Environment.load_hdr(str(p))
        assert (env.width, env.height) == (w, h)
        assert env.rgb.tobytes() == expect.tobytes()
    cv2 = pytest.importorskip("cv2")
    img = cv2.imread(str(rle), cv2.IMREAD_UNCHANGED)  # BGR, same RGBE formula
    if img is not None:
        assert np.allclose(img[..., ::-1], expect, rtol=1e-6, atol=0)
    with pytest.raises(rt.RtError):
        bad = tmp_path / "bad.hdr"
        bad.write_bytes(b"not an hdr")
        rt.Environment.load_hdr(str(bad))


def test_synthetic_hdri_is_rgbe_exact_and_deterministic(rt, tmp_path):
    a = rt.Environment.synthetic(128, 64)
    b = rt.Environment.synthetic(128, 64)
    assert a.rgb.tobytes() == b.rgb.tobytes()
    p = tmp_path / "sky.hdr"
    st = rt._ffi.host_lib().rt_hdr_write(str(p).encode(), 128, 64,
                                         a.rgb.ctypes.data_as(C.POINTER(C.c_float)))
    assert st == 0
    back = rt.Environment.load_hdr(str(p))
    assert back.rgb.tobytes() == a.rgb.tobytes()  # already quantised: round trip is exact
    assert a.rgb.max() > 1000 and a.rgb.min() > 0  # a sun and no black texels
    assert (a.rgba[..., 3] == 0).all() and a.rgba[..., :3].tobytes() == a.rgb.tobytes()


# ------------------------------------------------------------------ alias table
@pytest.mark.parametrize("w,h", [(64, 32), (256, 128)])
def test_alias_table_bit_exact_vs_numpy_restatement(rt, scene_oracle, w, h):
    env = rt.Environment.synthetic(w, h)
    p, a, pmf, leftover = scene_oracle.alias_table(env.rgb)
    got = env.alias
    assert got[:, 0].view(np.float32).tobytes() == p.tobytes()
    assert got[:, 1].tobytes() == a.tobytes()
    assert got[:, 2].view(np.float32).tobytes() == pmf.tobytes()
    assert (got[:, 3] == 0).all()
    assert env.leftover_count == leftover


def _alias(rt, fn, rgb):
    import ctypes as C
    rgb = np.ascontiguousarray(rgb, dtype=np.float32)
    h, w, _ = rgb.shape
    table = np.zeros((w * h, 4), dtype=np.uint32)
    leftover = C.c_uint32()
    st = fn(w, h, rgb.ctypes.data_as(C.POINTER(C.c_float)), table.ctypes.data_as(C.c_void_p), C.byref(leftover))
    assert st == 0
    return table, leftover.value


@pytest.mark.parametrize("case", ["sky2k", "random", "uniform", "tiny", "one_bright", "heavy_tail", "all_black_but_one"])
def test_threaded_stack_free_alias_build_equals_the_literal_loop(rt, env_2k, case, monkeypatch):
    # rt_alias_build (all host threads, cursors instead of stacks) vs rt_alias_build_serial (the
    # literal environments.rs:97-186 loop), byte for byte, also with other thread counts.
    rng = np.random.default_rng(7)
    rgb = {"sky2k": lambda: env_2k.environments[0].rgb,
           "random": lambda: rng.random((300, 517, 3), dtype=np.float32) ** 8 * 50,
           "uniform": lambda: np.full((64, 128, 3), 0.5, dtype=np.float32),
           "tiny": lambda: rng.random((3, 5, 3), dtype=np.float32),
           "one_bright": lambda: np.where(np.arange(64 * 32 * 3).reshape(32, 64, 3) // 3 == 777, 1000.0, 0.01),
           "heavy_tail": lambda: rng.random((512, 1024, 3), dtype=np.float32) ** 20 * 1e4,
           "all_black_but_one": lambda: np.where(np.arange(16 * 8 * 3).reshape(8, 16, 3) // 3 == 5, 3.0, 0.0)}[case]()
    lib = rt._ffi.host_lib()
    want, want_left = _alias(rt, lib.rt_alias_build_serial, rgb)
    for threads in ("", "1", "3"):
        if threads:
            monkeypatch.setenv("RT_HOST_THREADS", threads)
        got, got_left = _alias(rt, lib.rt_alias_build, rgb)
        assert got.tobytes() == want.tobytes() and got_left == want_left, f"threads={threads or 'all'}"


def test_alias_table_properties_full_size(env_2k):
    env = env_2k.environments[0]
    n = env.width * env.height
    prob = env.alias[:, 0].view(np.float32)
    alias = env.alias[:, 1]
    pmf = env.alias[:, 2].view(np.float32)
    assert len(prob) == n == 2048 * 1024
    assert (prob >= 0).all() and (prob <= 1.0).all() and (alias < n).all()
    # The table reproduces the pmf: P(i) = (prob_i + sum_{j: alias_j = i} (1 - prob_j)) / n.
    recon = prob.astype(np.float64).copy()
    np.add.at(recon, alias, 1.0 - prob.astype(np.float64))
    recon /= n
    assigned = ~((prob == 1.0) & (alias == np.arange(n)))
    rgb = env.rgb
    lum = ((np.float32(0.2126) * rgb[..., 0] + np.float32(0.7152) * rgb[..., 1]) +
           np.float32(0.0722) * rgb[..., 2])
    rows = np.float32(np.pi) * (np.arange(env.height, dtype=np.float32) + np.float32(0.5)) / np.float32(env.height)
    w = (lum * np.sin(rows.astype(np.float64)).astype(np.float32)[:, None]).reshape(-1)
    # The reference normalises by a sequential f32 sum (environments.rs:110), which on 2M
    # texels drifts ~0.1 % from the exact sum; the pmf inherits that scale factor.
    seq_sum = np.float64(np.cumsum(w, dtype=np.float32)[-1])
    assert abs(seq_sum / w.astype(np.float64).sum() - 1.0) < 5e-3
    true_pmf = w.astype(np.float64) / seq_sum
    # Entries the reference's loop never reaches (one stack runs dry first because of that
    # drift) keep probability 1 -> mass 1/N instead of their pmf.
    assert 0 < env.leftover_count < 4000
    assert abs(recon.sum() - 1.0) < 5e-3
    assert np.allclose(recon[assigned], true_pmf[assigned], rtol=1e-4, atol=1e-12)
    assert np.allclose(pmf[assigned], true_pmf[assigned], rtol=2e-5, atol=0)
    # Left-over entries carry the 1/N default (environments.rs:162-178).
    assert np.all(pmf[~assigned] == np.float32(1.0) / np.float32(n))
    assert env.leftover_count == int((~assigned).sum())
    # The sun is tiny but holds a large share of the probability mass.
    assert true_pmf[lum.reshape(-1) > 1000].sum() > 0.3


def test_image_writers_from_python(rt, tmp_path):
    # PFM: header, bottom row first, RGB of an RGBA buffer; PNG: signature and size fields.
    h, w = 5, 7
    rgba = np.arange(h * w * 4, dtype=np.float32).reshape(h, w, 4)
    pfm = tmp_path / "a.pfm"
    rt.write_pfm(str(pfm), rgba)
    raw = pfm.read_bytes()
    header = b"PF\n7 5\n-1.0\n"
    assert raw.startswith(header) and len(raw) == len(header) + h * w * 12
    data = np.frombuffer(raw[len(header):], dtype="<f4").reshape(h, w, 3)
    assert np.array_equal(data, rgba[::-1, :, :3])
    png = tmp_path / "a.png"
    rt.write_png(str(png), (np.arange(h * w * 3) % 256).astype(np.uint8).reshape(h, w, 3))
    raw = png.read_bytes()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n" and struct.unpack(">II", raw[16:24]) == (w, h)


def test_wave_kernels_keep_their_state_in_registers(rt):
    # Build hygiene read off the ptxas log of the in-tree build: an out-of-line helper with
    # reference parameters once pinned the best hit of the walk / pre-pass / finish kernels to
    # local memory (a fifth of the walk's load/store requests). Stack frames: none for pre-pass
    # and finish, spills only for shade and the shallow walk, the 64-entry stack for the deep walk.
    log = os.path.join(ROOT, "rsoderh-raytracing_b200", "csrc", "ptxas.log")
    if not os.path.exists(log):
        pytest.skip("no ptxas log (library not built in this tree)")
    text = open(log).read()
    frames = {}
    for m in re.finditer(r"Function properties for (\S+)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores", text):
        frames[m.group(1)] = (int(m.group(2)), int(m.group(3)))
    def frame(part):
        hits = [v for k, v in frames.items() if part in k]
        assert hits, f"{part} not in the ptxas log"
        return hits[0]
    assert frame("wave_prepass_kernel")[0] == 0
    assert frame("wave_finish_kernel")[0] <= 16
    assert frame("wave_walk_kernelILb1")[0] <= 32          # stack in shared memory: spills only
    assert frame("wave_walk_kernelILb0")[0] <= 512 + 32    # 64 x 8-byte stack entries + spills
    assert frame("wave_shade_kernel")[0] <= 64


def test_rust_declarations_are_generated_from_the_headers():
    """bindings/rust/src/rt_sys.rs is what tools/gen_rust_bindings.py writes from include/*.h today, and
    it declares every function of both headers with one parameter per C parameter."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_bindings.py"), "--check"])
    assert r.returncode == 0, "bindings/rust/src/rt_sys.rs is stale: run tools/gen_rust_bindings.py"
    rust = open(os.path.join(ROOT, "bindings", "rust", "src", "rt_sys.rs")).read()
    declared = dict(re.findall(r"pub fn (rt_\w+)\((.*?)\)", rust))
    for header in ("rt_backend.h", "rt_host.h"):
        text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", header)).read(), flags=re.S)
        for name, args in re.findall(r"\b(rt_\w+)\s*\(([^;{]*?)\)\s*;", text):
            assert name in declared, f"{name} is missing from rt_sys.rs"
            want = 0 if args.strip() == "void" else args.count(",") + 1
            got = 0 if not declared[name].strip() else declared[name].count(",") + 1
            assert got == want, (name, declared[name])
    for struct, size_fields in (("RtBvhNode", 8), ("RtSceneDesc", 16), ("RtCounters", 12)):
        body = re.search(r"pub struct %s \{(.*?)\n\}" % struct, rust, flags=re.S).group(1)
        assert body.count("pub ") == size_fields, struct
