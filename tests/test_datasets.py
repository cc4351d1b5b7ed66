"""CPU tests of the PhysicEdit training-data format (physicedit_b200/datasets.py) on a synthetic clip tree: selection rules, rule
splitting, frame-count / resolution / key-frame rules, and -- when a reference tree is around -- record-for-record and pixel-for-pixel
equality with the reference's own PhysicalEditingDataset (DiffSynth-Studio/diffsynth/trainers/utils.py:369-683) reading the same tree
through the same decoder."""
import json
import os
import sys
import types
import warnings

import numpy as np
import pytest
import torch

cv2 = pytest.importorskip("cv2")

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from physicedit_b200 import datasets as D  # noqa: E402


def write_clip(path, n_frames, w, h, seed):
    wr = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"mp4v"), 8, (w, h))
    assert wr.isOpened()
    yy, xx = np.mgrid[0:h, 0:w]
    for f in range(n_frames):
        img = np.stack([(xx * 2 + f * 5 + seed) % 256, (yy * 3 + f * 3) % 256, (xx + yy + f * 7 + seed * 11) % 256], axis=-1).astype(np.uint8)
        wr.write(img)
    wr.release()


def meta(idx, **over):
    rec = {"idx": idx, "prompt": f"original {idx}", "state": f"state {idx}", "transition": f"transition {idx}", "edit_instruction": f"edit {idx}",
           "triplet": {"subject": "ball", "idx": idx},
           "stage_a": {"principles": [
               {"id": "r_grav", "priority": "High", "instruction": " things fall ", "visual_cues": [" moves down ", "", 3], "negations": None},
               {"id": "r_low", "priority": "low", "instruction": "ignored"},
               {"priority": "HIGH", "instruction": "no id"},                       # -> rule_2
               "not a dict",                                                          # malformed: skipped
               {"id": "r_contra", "priority": "high", "instruction": "stays rigid"},
               {"id": "r_unknown", "priority": "high", "instruction": "no verdict"}]},
           "stage_b": {"rule_checks": [{"id": "r_grav", "result": "Supported", "matched_cues": ["moves down"]},
                                       {"id": "r_contra", "result": "contradicted"}, {"id": "rule_2", "result": "SUPPORTED"},
                                       {"id": "r_low", "result": "supported"}]}}
    rec.update(over)
    return rec


@pytest.fixture(scope="module")
def clip_tree(tmp_path_factory):
    root = tmp_path_factory.mktemp("clips")
    a, b = root / "scene_b" / "take1", root / "scene_a"
    (a / "deeper").mkdir(parents=True)
    b.mkdir(parents=True)
    write_clip(a / "0.mp4", 52, 96, 64, 1)            # >= 49 frames: the full window
    write_clip(a / "1.mp4", 10, 80, 48, 2)            # short clip: 9 frames (9 % 4 == 1)
    write_clip(a / "7.mp4", 12, 64, 64, 3)            # listed in the exclusion file
    write_clip(a / "9.mp4", 12, 64, 64, 4)            # no metadata record
    write_clip(a / "take.mp4", 12, 64, 64, 5)         # stem is not an integer
    write_clip(a / "deeper" / "3.mp4", 12, 64, 64, 6)  # below a clip directory: never visited
    write_clip(b / "2.mp4", 30, 200, 120, 7)
    lines = [json.dumps(meta(0, prompt="superseded")), "", "{broken json", json.dumps({"no": "idx"}), json.dumps(meta(0)), json.dumps(meta(1)),
             json.dumps(meta(7)), json.dumps(meta(3))]
    (a / D.METADATA_FILE).write_text("\n".join(lines) + "\n", encoding="utf-8")
    (a / D.EXCLUDED_FILE).write_text("7.mp4\n\n", encoding="utf-8")
    (a / "deeper" / D.METADATA_FILE).write_text(json.dumps(meta(3)) + "\n", encoding="utf-8")
    (b / D.METADATA_FILE).write_text(json.dumps(meta(2, stage_b={})) + "\n", encoding="utf-8")
    return root


def test_index_selection_and_rules(clip_tree, capsys):
    ds = D.PhysicalEditingDataset(root_dir=str(clip_tree), num_frames=49, height=48, width=80, repeat=3)
    assert "collected 3 samples from 2 leaf dirs" in capsys.readouterr().out
    assert [(os.path.basename(os.path.dirname(s["path"])), s["idx"]) for s in ds.samples] == [("scene_a", 2), ("take1", 0), ("take1", 1)]
    assert len(ds) == 9 and not ds.dynamic_resolution
    s0 = ds.samples[1]
    assert s0["original_prompt"] == "original 0" and s0["prompt"] == "edit 0" and s0["triplet"] == {"subject": "ball", "idx": 0}   # the later line won
    assert s0["supported_rules"] == [{"id": "r_grav", "instruction": "things fall", "matched_cues": ["moves down"]},
                                     {"id": "rule_2", "instruction": "no id", "matched_cues": []}]
    assert s0["contradicted_rules"] == [{"id": "r_contra", "instruction": "stays rigid"}]
    assert ds.samples[0]["supported_rules"] == [] and ds.samples[0]["contradicted_rules"] == []                       # no stage-B verdicts
    rules = D.high_priority_rules(meta(5))
    assert [r["id"] for r in rules] == ["r_grav", "rule_2", "r_contra", "r_unknown"]
    assert rules[0]["visual_cues"] == ["moves down", "3"] and rules[0]["negations"] == []
    with pytest.raises(TypeError):                                  # a record without stage_a is an error in the reference too (:473)
        D.high_priority_rules({"prompt": "x"})
    with pytest.raises(TypeError):
        D.PhysicalEditingDataset(root_dir=str(clip_tree), require_meta=False)      # 9.mp4 has no record: the default record has no stage_a


def test_samples_frames_and_key_frames(clip_tree):
    ds = D.PhysicalEditingDataset(root_dir=str(clip_tree), num_frames=49, height=48, width=80)
    with warnings.catch_warnings():
        warnings.simplefilter("error")                             # a 49-frame window gives exactly six key frames: no warning
        s = ds[1]
    assert set(s) == {"image", "edit_image", "middle_key_frames", "stitched_image", "prompt", "state", "transition", "idx", "path", "original_prompt",
                      "triplet", "supported_rules", "contradicted_rules"}
    assert s["idx"] == 0 and s["image"].size == (80, 48) and s["edit_image"].size == (80, 48)
    assert len(s["middle_key_frames"]) == 6 and s["stitched_image"].size == (160, 144)
    assert np.array_equal(np.asarray(s["stitched_image"].crop((80, 48, 160, 96))), np.asarray(s["middle_key_frames"][3]))      # row 1, column 1
    assert not np.array_equal(np.asarray(s["image"]), np.asarray(s["edit_image"]))
    with pytest.warns(UserWarning, match="Expected 6 frames, but got 1"):
        short = ds[2]                                              # 10-frame clip -> 9 frames -> 7 inner frames -> one run -> one key frame
    assert len(short["middle_key_frames"]) == 1 and short["stitched_image"] is None
    assert ds[2 + 3]["idx"] == ds[2]["idx"]                        # repeat wraps around
    # dynamic resolution: frames keep their own (floored) size, scaled down to max_pixels
    dyn = D.PhysicalEditingDataset(root_dir=str(clip_tree), num_frames=5, max_pixels=100 * 62)
    assert dyn.dynamic_resolution and dyn[0]["image"].size == (96, 48)            # 200 x 120 -> 101 x 61 -> floored to multiples of 16
    assert dyn[1]["image"].size == (96, 64)                                        # 96 x 64 fits: unchanged
    assert len(dyn[1]["middle_key_frames"]) == 1


def test_frame_count_rule_and_geometry():
    ds = D.PhysicalEditingDataset.__new__(D.PhysicalEditingDataset)
    ds.num_frames, ds.time_division_factor, ds.time_division_remainder = 81, 4, 1

    class Src:
        def __init__(self, n): self.n = n
        def count(self): return self.n
    assert [ds._get_num_frames(Src(n)) for n in (200, 81, 80, 78, 6, 5, 4, 2, 1, 0)] == [81, 81, 77, 77, 5, 5, 1, 1, 1, 1]
    from PIL import Image
    img = Image.fromarray((np.arange(60 * 100 * 3) % 256).astype(np.uint8).reshape(60, 100, 3))
    out = D.cover_and_center_crop(img, 32, 32)
    assert out.size == (32, 32)
    frames = [Image.new("RGB", (4, 4), (i, 0, 0)) for i in range(20)]
    picked = D.middle_key_frames(frames, 8)                          # inner = frames 1..18 -> runs [1..8] [9..16] [17, 18] -> centres 5, 13, 18
    assert [p.getpixel((0, 0))[0] for p in picked] == [5, 13, 18]
    assert D.middle_key_frames(frames[:2], 8) == []


def test_train_script_names_resolve_to_the_dataset(clip_tree):
    from physicedit_b200 import compat
    saved = {k: v for k, v in sys.modules.items() if k == "diffsynth" or k.startswith("diffsynth.")}
    try:
        compat.install()
        from diffsynth.trainers.utils import PhysicalEditingDataset, qwen_image_parser
        args = qwen_image_parser().parse_args(["--dataset_base_path", str(clip_tree), "--dinov2_path", "unused", "--height", "48", "--width", "80",
                                               "--num_frames", "49", "--dataset_repeat", "2"])
        ds = PhysicalEditingDataset(args=args)                     # scripts/train/train_physicedit.py:420
        assert isinstance(ds, D.PhysicalEditingDataset) and len(ds) == 6 and ds.num_frames == 49 and (ds.height, ds.width) == (48, 80)
        loader = torch.utils.data.DataLoader(ds, shuffle=False, collate_fn=lambda x: x[0], num_workers=0)      # as launch_training_task builds it
        first = next(iter(loader))
        assert first["idx"] == 2 and first["image"].size == (80, 48)
    finally:
        for k in [k for k in sys.modules if k == "diffsynth" or k.startswith("diffsynth.")]:
            del sys.modules[k]
        sys.modules.update(saved)


# ---- against the reference class itself -------------------------------------------------------------------------------------------------
class _Cv2Reader:
    """What the reference asks of `imageio.get_reader(path)` (count_frames / get_data / close), served by the decoder the product uses here."""

    def __init__(self, path):
        self.src = D._OpenCVFrames(path)

    def count_frames(self): return self.src.count()
    def get_data(self, i): return self.src.frame(i)
    def close(self): self.src.close()


def _import_reference_dataset():
    """The reference's class, unmodified, compiled from its own source file: only the `VIDEO_EXTS` constant and the `PhysicalEditingDataset`
    class statement of trainers/utils.py are executed (the module's other imports -- peft, accelerate, imageio -- are absent here), with
    `imageio.get_reader` served by the OpenCV reader."""
    import ast
    import typing
    from pathlib import Path
    import torchvision
    from PIL import Image
    from oracle.ref_import import reference_root
    root = reference_root()
    if root is None:
        pytest.skip("no reference tree (baseline/_ref or /root/reference)")
    path = os.path.join(root, "trainers", "utils.py")
    tree = ast.parse(open(path, encoding="utf-8").read())
    keep = [n for n in tree.body if (isinstance(n, ast.ClassDef) and n.name == "PhysicalEditingDataset")
            or (isinstance(n, ast.Assign) and any(getattr(t, "id", None) == "VIDEO_EXTS" for t in n.targets))]
    assert len(keep) == 2
    imageio = types.ModuleType("imageio")
    imageio.get_reader = _Cv2Reader
    ns = dict(imageio=imageio, os=os, torch=torch, warnings=warnings, torchvision=torchvision, json=json, Image=Image, Path=Path,
              Optional=typing.Optional, List=typing.List, Dict=typing.Dict, Any=typing.Any, Set=typing.Set, Tuple=typing.Tuple)
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return ns["PhysicalEditingDataset"]


def test_matches_the_reference_dataset_record_for_record_and_pixel_for_pixel(clip_tree, monkeypatch):
    Ref = _import_reference_dataset()
    monkeypatch.setattr(D, "open_video", lambda p: D._OpenCVFrames(p))             # same decoder on both sides
    for kw in (dict(num_frames=49, height=48, width=80), dict(num_frames=13, max_pixels=100 * 60), dict(num_frames=81, height=64, width=64, key_frame_stride=3)):
        ref = Ref(root_dir=str(clip_tree), **kw)
        ours = D.PhysicalEditingDataset(root_dir=str(clip_tree), **kw)
        assert ours.samples == ref.samples and len(ours) == len(ref) and ours.dynamic_resolution == ref.dynamic_resolution
        for i in range(len(ref)):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                r, o = ref[i], ours[i]
            assert set(r) == set(o)
            for k in r:
                if k in ("image", "edit_image"):
                    assert o[k].size == r[k].size and np.array_equal(np.asarray(o[k]), np.asarray(r[k])), (kw, i, k)
                elif k == "middle_key_frames":
                    assert len(o[k]) == len(r[k]) and all(np.array_equal(np.asarray(a), np.asarray(b)) for a, b in zip(o[k], r[k])), (kw, i)
                elif k == "stitched_image":
                    assert (o[k] is None) == (r[k] is None) and (o[k] is None or np.array_equal(np.asarray(o[k]), np.asarray(r[k]))), (kw, i)
                else:
                    assert o[k] == r[k], (kw, i, k)


def test_pica100k_dataset_matches_the_reference(monkeypatch):
    """Pica100kDataset (trainers/utils.py:685-775) over in-memory records standing in for the Hugging Face download: same samples as the reference
    class, pixel for pixel, fixed and dynamic resolution; a record without images warns and yields None."""
    import ast
    import typing
    import datasets as hf
    import torchvision
    from PIL import Image
    from oracle.ref_import import reference_root
    g = np.random.default_rng(0)
    pic = lambda w, h: Image.fromarray(g.integers(0, 255, (h, w, 3), dtype=np.uint8))
    records = [dict(src_img=pic(300, 200), tgt_img=pic(310, 190).convert("L"), superficial_prompt="make it rain"), dict(src_img=pic(64, 64), tgt_img=None),
               dict(src_img=pic(1000, 700), tgt_img=pic(1000, 700))]
    seen = {}

    def load_dataset(name, split=None, cache_dir=None):
        seen.update(name=name, split=split, cache_dir=cache_dir)
        return records
    monkeypatch.setattr(hf, "load_dataset", load_dataset)
    ours = [D.Pica100kDataset(cache_dir="/data", height=128, width=192, repeat=2), D.Pica100kDataset(max_pixels=400 * 300)]
    assert seen == dict(name="Andrew613/PICA-100K", split="train", cache_dir="/data") or seen["cache_dir"] is None
    assert len(ours[0]) == 6 and not ours[0].dynamic_resolution and ours[1].dynamic_resolution
    s = ours[0][3]
    assert set(s) == {"image", "edit_image", "prompt"} and s["prompt"] == "make it rain" and s["image"].size == (192, 128) and s["image"].mode == "RGB"
    assert isinstance(s["edit_image"], list) and s["edit_image"][0].size == (192, 128)
    with pytest.warns(UserWarning, match="missing src_img/tgt_img"):
        assert ours[0][1] is None
    assert ours[1][2]["image"].size == (400, 288)                      # 1000 x 700 -> 414 x 289 -> floored to multiples of 16
    root = reference_root()
    if root is None:
        return
    path = os.path.join(root, "trainers", "utils.py")
    cls = [n for n in ast.parse(open(path, encoding="utf-8").read()).body if isinstance(n, ast.ClassDef) and n.name == "Pica100kDataset"]
    ns = dict(torch=torch, warnings=warnings, torchvision=torchvision, Image=Image, load_dataset=load_dataset, Optional=typing.Optional, Dict=typing.Dict,
              Any=typing.Any, Tuple=typing.Tuple)
    exec(compile(ast.Module(body=cls, type_ignores=[]), path, "exec"), ns)
    refs = [ns["Pica100kDataset"](cache_dir="/data", height=128, width=192, repeat=2), ns["Pica100kDataset"](max_pixels=400 * 300)]
    for o, r in zip(ours, refs):
        assert len(o) == len(r)
        for i in (0, 2):
            a, b = o[i], r[i]
            assert a["prompt"] == b["prompt"] and np.array_equal(np.asarray(a["image"]), np.asarray(b["image"]))
            assert len(a["edit_image"]) == len(b["edit_image"]) == 1 and np.array_equal(np.asarray(a["edit_image"][0]), np.asarray(b["edit_image"][0]))


def test_unified_dataset_and_operators_match_the_reference(tmp_path, clip_tree):
    """`diffsynth.trainers.unified_dataset` (the whole reference module executed from its source, `imageio` served by the OpenCV reader) against
    physicedit_b200/unified_dataset.py: metadata in all three formats, the default image / video operators (still image, clip, list of images),
    operator composition with `>>`, routing errors, and the cache mode that reads what `launch_data_process_task` writes."""
    import types
    import pandas
    from PIL import Image
    from oracle.ref_import import reference_root
    from physicedit_b200 import unified_dataset as U
    g = np.random.default_rng(1)
    base = tmp_path / "data"
    (base / "imgs").mkdir(parents=True)
    for name, (w, h) in (("a.png", (300, 200)), ("b.jpg", (640, 640)), ("c.webp", (100, 180))):
        Image.fromarray(g.integers(0, 255, (h, w, 3), dtype=np.uint8)).save(base / "imgs" / name)
    write_clip(base / "clip.mp4", 11, 96, 64, 4)
    rows = [{"image": "imgs/a.png", "prompt": "one", "score": 1.5}, {"image": ["imgs/b.jpg", "imgs/c.webp"], "prompt": "two", "score": 2}]
    (base / "meta.json").write_text(json.dumps(rows))
    (base / "meta.jsonl").write_text("".join(json.dumps(r) + "\n" for r in rows))
    pandas.DataFrame([{"image": "imgs/a.png", "prompt": "one"}, {"image": "imgs/b.jpg", "prompt": "two"}]).to_csv(base / "meta.csv", index=False)
    cache = tmp_path / "cache"
    (cache / "0").mkdir(parents=True)
    (cache / "1").mkdir()
    torch.save({"prompt_emb": torch.arange(6.0).view(1, 2, 3), "height": 64}, cache / "0" / "0.pth")
    torch.save({"prompt_emb": torch.ones(1, 2, 3), "height": 96}, cache / "1" / "0.pth")
    (cache / "1" / "notes.txt").write_text("ignored")

    def scenarios(M):
        out = {}
        for meta_name in ("meta.json", "meta.jsonl", "meta.csv"):
            ds = M.UnifiedDataset(base_path=str(base), metadata_path=str(base / meta_name), repeat=2, data_file_keys=("image",),
                                  main_data_operator=M.UnifiedDataset.default_image_operator(base_path=str(base), max_pixels=200 * 200, height_division_factor=16,
                                                                                             width_division_factor=16))
            out[meta_name] = (len(ds), ds.load_from_cache, [ds[i] for i in range(len(ds))])
        fixed = M.UnifiedDataset(base_path=str(base), metadata_path=str(base / "meta.json"), data_file_keys=("image", "score"),
                                 main_data_operator=M.UnifiedDataset.default_image_operator(base_path=str(base), height=96, width=128),
                                 special_operator_map={"score": M.ToFloat() >> M.ToStr() >> M.ToList()})
        out["fixed"] = [fixed[0], fixed[1]]
        video = M.UnifiedDataset.default_video_operator(base_path=str(base), height=48, width=80, num_frames=9)
        out["video"] = [video("clip.mp4"), video("imgs/a.png")]
        short = M.UnifiedDataset.default_video_operator(base_path=str(base), num_frames=81, max_pixels=64 * 48)
        out["short clip"] = short("clip.mp4")                                  # 11 frames -> 9 (9 % 4 == 1), dynamic resolution
        cached = M.UnifiedDataset(base_path=str(cache), repeat=3)
        out["cache"] = (len(cached), cached.load_from_cache, sorted((d["height"], d["prompt_emb"].sum().item()) for d in (cached[0], cached[1])))
        chain = M.ToAbsolutePath("/x") >> (M.ToStr() >> M.ToList())
        out["chain"] = (chain("y"), len(chain.operators), M.ToStr(none_value="n/a")(None), M.ToInt()("7"), M.DataProcessingOperatorRaw()(3))
        for bad in (lambda: M.RouteByExtensionName([(("png",), M.LoadImage())])("x.tiff"), lambda: M.RouteByType([(str, M.ToStr())])(3),
                    lambda: M.DataProcessingOperator()(1)):
            with pytest.raises((ValueError, NotImplementedError)):
                bad()
        return out

    def equal(a, b, path="out"):
        if isinstance(a, Image.Image):
            assert isinstance(b, Image.Image) and a.size == b.size and a.mode == b.mode and a.tobytes() == b.tobytes(), path
        elif isinstance(a, dict):
            assert set(a) == set(b), path
            for k in a:
                equal(a[k], b[k], f"{path}.{k}")
        elif isinstance(a, (list, tuple)):
            assert len(a) == len(b), path
            for i, (x, y) in enumerate(zip(a, b)):
                equal(x, y, f"{path}[{i}]")
        elif isinstance(a, float) and a != a:
            assert b != b, path
        else:
            assert a == b, (path, a, b)

    ours = scenarios(U)
    assert ours["meta.json"][0] == 4 and not ours["meta.json"][1] and ours["meta.json"][2][0]["image"].size == (240, 160)       # 300 x 200 -> 244 x 163 -> / 16
    assert [im.size for im in ours["meta.json"][2][1]["image"]] == [(192, 192), (96, 176)]
    assert ours["fixed"][0]["image"].size == (128, 96) and ours["fixed"][0]["score"] == ["1.5"]
    assert len(ours["video"][0]) == 9 and ours["video"][0][0].size == (80, 48) and len(ours["video"][1]) == 1 and len(ours["short clip"]) == 9
    assert ours["cache"] == (6, True, [(64, 15.0), (96, 6.0)]) and ours["chain"][:2] == (["/x/y"], 3)
    root = reference_root()
    if root is None:
        return
    path = os.path.join(root, "trainers", "unified_dataset.py")
    imageio = types.ModuleType("imageio")
    imageio.get_reader = _Cv2Reader
    iio = types.ModuleType("imageio.v3")
    saved = {k: sys.modules.get(k) for k in ("imageio", "imageio.v3")}
    sys.modules["imageio"], sys.modules["imageio.v3"] = imageio, iio
    try:
        ns = {"__name__": "ref_unified_dataset"}
        exec(compile(open(path, encoding="utf-8").read(), path, "exec"), ns)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    import unittest.mock
    with unittest.mock.patch.object(D, "open_video", lambda p: D._OpenCVFrames(p)), unittest.mock.patch.object(U, "open_video", lambda p: D._OpenCVFrames(p)):
        equal(scenarios(types.SimpleNamespace(**ns)), scenarios(U))


def test_open_video_falls_back_and_reports(tmp_path, monkeypatch):
    clip = tmp_path / "c.mp4"
    write_clip(clip, 5, 64, 48, 1)
    src = D.open_video(str(clip))                      # imageio is absent here (or cannot open it): OpenCV takes over
    assert src.count() == 5 and src.frame(0).shape == (48, 64, 3)
    src.close()
    with pytest.raises(OSError, match="cannot open"):
        D.open_video(str(tmp_path / "missing.mp4"))
    ds = D.PhysicalEditingDataset.__new__(D.PhysicalEditingDataset)
    ds.num_frames, ds.time_division_factor, ds.time_division_remainder = 81, 4, 1
    with pytest.warns(UserWarning, match="cannot open video"):
        assert ds._load_video(str(tmp_path / "missing.mp4")) == []


def test_only_the_frames_a_sample_uses_are_resized(clip_tree):
    ds = D.PhysicalEditingDataset(root_dir=str(clip_tree), num_frames=49, height=48, width=80)
    path = ds.samples[1]["path"]
    full, lazy = ds._load_video(path), ds._load_video(path, only_used=True)
    assert len(full) == len(lazy) == 49 and all(f is not None for f in full)
    used = [i for i, f in enumerate(lazy) if f is not None]
    assert used == [0, 5, 13, 21, 29, 37, 44, 48]                    # first, the six run centres of frames 1..47 (the last run has 7 frames), last
    assert all(np.array_equal(np.asarray(lazy[i]), np.asarray(full[i])) for i in used)


def test_a_clip_shorter_than_its_header_says(monkeypatch):
    """The header promises 20 frames, the stream ends after 12: the sample is built from the 12 that exist (first, last = frame 11, key frames of 1..10),
    exactly what resizing everything and picking afterwards gives."""
    class Truncated:
        def __init__(self, path): self.closed = False
        def count(self): return 20
        def frame(self, i):
            if i >= 12: raise IndexError(i)
            return np.full((32, 48, 3), i * 10, np.uint8)
        def skip(self, i): return i < 12
        def close(self): self.closed = True
    monkeypatch.setattr(D, "open_video", Truncated)
    ds = D.PhysicalEditingDataset.__new__(D.PhysicalEditingDataset)
    ds.num_frames, ds.time_division_factor, ds.time_division_remainder, ds.key_frame_stride = 81, 4, 1, 4
    ds.dynamic_resolution, ds.height, ds.width = False, 32, 48
    full, lazy = ds._load_video("x.mp4"), ds._load_video("x.mp4", only_used=True)
    assert len(full) == len(lazy) == 12                                  # 20 -> 17 by the frame-count rule, 12 readable
    used = [i for i, f in enumerate(lazy) if f is not None]
    assert used == [0, 3, 7, 10, 11] and all(lazy[i].getpixel((0, 0)) == full[i].getpixel((0, 0)) == (i * 10,) * 3 for i in used)
    assert [f.getpixel((0, 0))[0] for f in ds.extract_middle_key_frames(lazy)] == [30, 70, 100]
