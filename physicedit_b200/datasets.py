"""The PhysicEdit training-data formats: `PhysicalEditingDataset`, the one dataset scripts/train/train_physicedit.py builds (:420) and whose
sample dictionaries feed `QwenImageTrainingModule.forward_preprocess` (:255-296) and through it the pipeline's units; and `Pica100kDataset`
(image pairs of PICA-100K, trainers/utils.py:685-775).

Mirrors DiffSynth-Studio/diffsynth/trainers/utils.py:367-775 (constructor signature, `samples` records, sample keys, frame-count /
resolution / key-frame rules, warnings and error behaviour); host-side only -- frames are decoded on the CPU and handed to the units as PIL
images exactly like the reference does, the GPU work starts at the VAE / DINOv2 encoders.

On-disk layout (one "clip directory" = any directory that directly holds video files; directories below it are not visited):
    <clip dir>/<integer>.mp4 ...              a clip: first frame = the edit (source) image, last frame = the target image
    <clip dir>/unified_output_new_qwen.jsonl  one JSON object per line, keyed by "idx" = the clip's integer stem
    <clip dir>/final_filter_videos.txt        optional: file names to leave out
Video decoding uses imageio when it is installed (the reference's decoder) and OpenCV otherwise; with neither, loading a clip raises.
"""
from __future__ import annotations

import json
import os
import warnings
from pathlib import Path
from typing import Any, Dict, List, Optional, Set, Tuple

import torch
from PIL import Image

VIDEO_EXTS = {".mp4", ".mov", ".mkv", ".avi", ".webm", ".m4v"}          # trainers/utils.py:15
METADATA_FILE = "unified_output_new_qwen.jsonl"                          # :437
EXCLUDED_FILE = "final_filter_videos.txt"                                # :460


# ---- frame sources ----------------------------------------------------------------------------------------------------------------
class _ImageioFrames:
    def __init__(self, path):
        import imageio
        self.reader = imageio.get_reader(path)

    def count(self) -> int:
        try:
            return int(self.reader.count_frames())
        except Exception:  # noqa: BLE001  (formats without a frame count: probe, :580-587)
            n = 0
            try:
                while True:
                    self.reader.get_data(n)
                    n += 1
            except Exception:  # noqa: BLE001
                return n

    def frame(self, i):
        return self.reader.get_data(i)

    def skip(self, i) -> bool:
        return True                                # random access: nothing to do for a frame nobody asks for

    def close(self):
        self.reader.close()


class _OpenCVFrames:
    def __init__(self, path):
        import cv2
        self.cv2 = cv2
        self.cap = cv2.VideoCapture(path)
        if not self.cap.isOpened():
            raise OSError(f"OpenCV cannot open {path}")
        self.next = 0

    def count(self) -> int:
        n = int(self.cap.get(self.cv2.CAP_PROP_FRAME_COUNT))
        if n > 0:
            return n
        n = 0                                      # containers without a frame count: walk the stream once, then rewind
        while self.cap.grab():
            n += 1
        self.cap.set(self.cv2.CAP_PROP_POS_FRAMES, 0)
        self.next = 0
        return n

    def frame(self, i):
        if i != self.next:
            self.cap.set(self.cv2.CAP_PROP_POS_FRAMES, i)
        ok, bgr = self.cap.read()
        if not ok:
            raise IndexError(f"no frame {i}")
        self.next = i + 1
        return self.cv2.cvtColor(bgr, self.cv2.COLOR_BGR2RGB)      # RGB, like imageio

    def skip(self, i) -> bool:
        """Advance past frame i without converting / copying it; False when the stream ends there."""
        if i != self.next:
            self.cap.set(self.cv2.CAP_PROP_POS_FRAMES, i)
        ok = self.cap.grab()
        self.next = i + 1
        return bool(ok)

    def close(self):
        self.cap.release()


def open_video(path: str):
    """A frame source with count() / frame(i) -> uint8 RGB array / close(): imageio (the reference's decoder) when it is installed and can open the
    file, else OpenCV; with neither installed an ImportError, with both failing the first decoder's error."""
    errors = []
    for backend in (_ImageioFrames, _OpenCVFrames):
        try:
            return backend(path)
        except ImportError as e:
            errors.append(e)
        except Exception as e:  # noqa: BLE001  (e.g. imageio without its ffmpeg plugin: let the other decoder try)
            errors.append(e)
    if all(isinstance(e, ImportError) for e in errors):
        raise ImportError("decoding training clips needs imageio (the reference's decoder) or OpenCV; neither is installed") from errors[0]
    raise next(e for e in errors if not isinstance(e, ImportError))


# ---- metadata ---------------------------------------------------------------------------------------------------------------------
def high_priority_rules(meta: Dict[str, Any]) -> List[Dict[str, Any]]:
    """The `priority: high` principles of a record's stage-A analysis (:472-491).  A record without `stage_a.principles` is an error, as in
    the reference (TypeError / KeyError from the lookup); a malformed principle is skipped."""
    rules = []
    for i, p in enumerate(meta.get("stage_a", [])["principles"]):
        try:
            if str(p.get("priority", "")).lower() != "high":
                continue
            clean = lambda xs: [str(x).strip() for x in (xs or []) if str(x).strip()]
            rules.append({"id": str(p.get("id") or f"rule_{i}"), "instruction": str(p.get("instruction", "")).strip(),
                          "visual_cues": clean(p.get("visual_cues", [])), "negations": clean(p.get("negations", []))})
        except Exception:  # noqa: BLE001
            continue
    return rules


def rule_outcomes(meta: Dict[str, Any], rules: List[Dict[str, Any]]) -> Tuple[List[Dict[str, Any]], List[Dict[str, Any]]]:
    """Splits the high-priority rules by the stage-B verdict on each (:493-512): (supported, contradicted); anything else is dropped."""
    verdicts = {rc.get("id", ""): rc for rc in meta.get("stage_b", {}).get("rule_checks", [])}
    supported, contradicted = [], []
    for r in rules:
        rid = r.get("id", "")
        rc = verdicts.get(rid, {})
        verdict = str(rc.get("result", "unknown")).lower()
        if verdict == "supported":
            supported.append({"id": rid, "instruction": r.get("instruction", ""), "matched_cues": rc.get("matched_cues", [])})
        elif verdict == "contradicted":
            contradicted.append({"id": rid, "instruction": r.get("instruction", "")})
    return supported, contradicted


# ---- frames -----------------------------------------------------------------------------------------------------------------------
def cover_and_center_crop(image: Image.Image, target_height: int, target_width: int) -> Image.Image:
    """Bilinear resize so that the image covers the target, then a centred crop (:551-560; torchvision's resize / center_crop on a PIL image)."""
    w, h = image.size
    scale = max(target_width / w, target_height / h)
    rh, rw = round(h * scale), round(w * scale)
    image = image.resize((rw, rh), Image.BILINEAR)
    top, left = int(round((rh - target_height) / 2.0)), int(round((rw - target_width) / 2.0))
    return image.crop((left, top, left + target_width, top + target_height))


def snapped_size(image: Image.Image, max_pixels: int, height_division_factor: int, width_division_factor: int) -> Tuple[int, int]:
    """Dynamic resolution (:562-574, :739-751): the image's own (height, width), scaled down to `max_pixels` and floored to the division factors."""
    width, height = image.size
    if width * height > max_pixels:
        scale = (width * height / max_pixels) ** 0.5
        height, width = int(height / scale), int(width / scale)
    snap = lambda v, f: max(f, v // f * f)
    return snap(height, height_division_factor), snap(width, width_division_factor)


def resolution_mode(height, width) -> bool:
    """True = dynamic resolution: unless BOTH sides are given (:406-414, :716-724, messages included)."""
    print({(False, False): "Height and width are fixed. Setting `dynamic_resolution` to False.",
           (True, True): "Height and width are none. Setting `dynamic_resolution` to True."}.get(
               (height is None, width is None), "One of height/width is None. Setting `dynamic_resolution` to True."))
    return height is None or width is None


def middle_key_frames(frames: List[Image.Image], stride: int) -> List[Image.Image]:
    """The centre frame of every `stride`-long run of the frames strictly between the first and the last (:620-633)."""
    inner = frames[1:-1] if len(frames) > 2 else []
    runs = [inner[i:i + stride] for i in range(0, len(inner), stride)]
    return [run[len(run) // 2] for run in runs if run]


def stitch_key_frames(frames: List[Image.Image]) -> Optional[Image.Image]:
    """Six key frames on a 2-wide, 3-high canvas of the first frame's cell size (:635-651); any other count warns and yields None."""
    if len(frames) != 6:
        warnings.warn(f"Expected 6 frames, but got {len(frames)}")
        return None
    w, h = frames[0].size
    canvas = Image.new("RGB", (2 * w, 3 * h))
    for i, img in enumerate(frames):
        canvas.paste(img, ((i % 2) * w, (i // 2) * h))
    return canvas


class PhysicalEditingDataset(torch.utils.data.Dataset):
    """trainers/utils.py:369-683.  `samples`: one record per usable clip (path, idx, prompts, rules); `__getitem__` decodes the clip."""

    def __init__(self, root_dir: str = None, num_frames: int = 81, time_division_factor: int = 4, time_division_remainder: int = 1,
                 max_pixels: int = 1920 * 1080, height: Optional[int] = None, width: Optional[int] = None, height_division_factor: int = 16,
                 width_division_factor: int = 16, video_file_extension=("mp4", "avi", "mov", "wmv", "mkv", "flv", "webm"), repeat: int = 1,
                 key_frame_stride: int = 8, require_meta: bool = True, args=None):
        if args is not None:                                # the train script passes its parsed flags (:420)
            root_dir = getattr(args, "dataset_base_path", root_dir)
            num_frames = getattr(args, "num_frames", num_frames)
            height, width = getattr(args, "height", height), getattr(args, "width", width)
            max_pixels = getattr(args, "max_pixels", max_pixels)
            repeat = getattr(args, "dataset_repeat", repeat)
        self.root = Path(root_dir)
        self.num_frames, self.repeat, self.key_frame_stride = int(num_frames), int(repeat), int(key_frame_stride)
        self.time_division_factor, self.time_division_remainder = int(time_division_factor), int(time_division_remainder)
        self.max_pixels, self.height, self.width = int(max_pixels), height, width
        self.height_division_factor, self.width_division_factor = int(height_division_factor), int(width_division_factor)
        self.require_meta = bool(require_meta)
        self.video_file_extension = video_file_extension
        self.dynamic_resolution = resolution_mode(self.height, self.width)
        self.samples: List[Dict[str, Any]] = self._build_samples(self.root)
        if not self.samples:
            warnings.warn("PhysicalEditingDataset: no valid samples found.")

    # -- index --------------------------------------------------------------------------------------------------------------------
    @staticmethod
    def _is_video_file(p: Path) -> bool:
        return p.suffix.lower() in VIDEO_EXTS

    def _collect_leaf_dirs(self, root: Path) -> List[Path]:
        found = []
        for cur, subdirs, files in os.walk(root):
            if any(self._is_video_file(Path(cur) / f) for f in files):
                found.append(Path(cur))
                subdirs[:] = []                            # a clip directory's children are not searched
        return sorted(set(found))

    @staticmethod
    def _read_leaf_metadata(leaf: Path) -> Dict[int, Dict[str, Any]]:
        records: Dict[int, Dict[str, Any]] = {}
        path = leaf / METADATA_FILE
        if path.exists():
            for line in path.read_text(encoding="utf-8").splitlines():
                try:
                    obj = json.loads(line)
                    records[int(obj["idx"])] = obj          # a later line with the same idx wins; unparsable lines are skipped
                except Exception:  # noqa: BLE001
                    continue
        return records

    @staticmethod
    def _read_filtered_names(leaf: Path) -> Set[str]:
        path = leaf / EXCLUDED_FILE
        return {n.strip() for n in path.read_text(encoding="utf-8").splitlines() if n.strip()} if path.exists() else set()

    def _list_videos(self, leaf: Path) -> List[Path]:
        return sorted(p for p in leaf.iterdir() if p.is_file() and self._is_video_file(p))

    read_high_rules = staticmethod(high_priority_rules)
    get_supported_and_contradicted_rules = staticmethod(rule_outcomes)

    def _build_samples(self, root: Path) -> List[Dict[str, Any]]:
        samples = []
        leaves = self._collect_leaf_dirs(root)
        for leaf in leaves:
            records, excluded = self._read_leaf_metadata(leaf), self._read_filtered_names(leaf)
            for clip in self._list_videos(leaf):
                if clip.name in excluded or not clip.stem.isdigit():
                    continue
                idx = int(clip.stem)
                meta = records.get(idx)
                if meta is None:
                    if self.require_meta:
                        continue
                    meta = {"prompt": "", "state": "", "transition": "", "edit_instruction": "", "triplet": {}}
                supported, contradicted = rule_outcomes(meta, high_priority_rules(meta))
                samples.append({"path": str(clip.resolve()), "idx": idx, "original_prompt": meta.get("prompt", ""), "state": meta.get("state", ""),
                                "transition": meta.get("transition", ""), "triplet": meta.get("triplet", {}), "prompt": meta.get("edit_instruction", ""),
                                "supported_rules": supported, "contradicted_rules": contradicted})
        samples.sort(key=lambda s: (Path(s["path"]).parent.as_posix(), s["idx"]))
        print(f"[PhysicalEditingDataset] collected {len(samples)} samples from {len(leaves)} leaf dirs.")
        return samples

    # -- clips --------------------------------------------------------------------------------------------------------------------
    _crop_and_resize = staticmethod(cover_and_center_crop)

    def _get_height_width(self, image: Image.Image) -> Tuple[int, int]:
        """Fixed (height, width), or -- dynamic resolution -- the frame's own size scaled down to `max_pixels` and floored to the division factors (:562-574)."""
        if not self.dynamic_resolution:
            return self.height, self.width
        return snapped_size(image, self.max_pixels, self.height_division_factor, self.width_division_factor)

    def _get_num_frames(self, source) -> int:
        """`num_frames`, or for a shorter clip the largest n <= its length with n % time_division_factor == time_division_remainder (:576-593)."""
        n = self.num_frames
        total = source.count()
        if total < n:
            n = total
            while n > 1 and n % self.time_division_factor != self.time_division_remainder:
                n -= 1
        return max(1, n)

    def _load_video(self, file_path: str, only_used: bool = False) -> List[Optional[Image.Image]]:
        """The clip's first `_get_num_frames` frames, RGB, cover-resized and centre-cropped (:595-618).  `only_used=True` (what `__getitem__` asks for)
        resizes only the frames a sample is made of -- first, last, the middle key frames: 8 of 49 -- and leaves None in the other slots; every frame
        is processed on its own, so those eight are the same images either way."""
        try:
            source = open_video(file_path)
        except ImportError:
            raise
        except Exception as e:  # noqa: BLE001
            warnings.warn(f"cannot open video {file_path}: {e}")
            return []
        try:
            n = self._get_num_frames(source)
            used = set(range(n)) if not only_used else {0, n - 1} | set(middle_key_frames(list(range(n)), self.key_frame_stride))
            frames: List[Optional[Image.Image]] = []
            for i in range(n):
                try:
                    if i not in used:
                        if not source.skip(i):
                            raise IndexError(f"no frame {i}")
                        frames.append(None)
                        continue
                    data = source.frame(i)
                except Exception:  # noqa: BLE001  (a clip shorter than its header says: keep what was read)
                    break
                img = Image.fromarray(data).convert("RGB")
                frames.append(cover_and_center_crop(img, *self._get_height_width(img)))
            if only_used and len(frames) < n:               # the clip ended early: which frames a sample uses depends on the real length
                full = self._load_video(file_path)
                m = len(full)
                keep = {0, m - 1} | set(middle_key_frames(list(range(m)), self.key_frame_stride))
                return [f if i in keep else None for i, f in enumerate(full)]
        except Exception as e:  # noqa: BLE001
            warnings.warn(f"error reading video {file_path}: {e}")
            frames = []
        finally:
            source.close()
        return frames

    def extract_middle_key_frames(self, frames: List[Image.Image]) -> List[Image.Image]:
        return middle_key_frames(frames, self.key_frame_stride)

    stitch_middle_key_frames = staticmethod(stitch_key_frames)

    def __len__(self) -> int:
        return len(self.samples) * self.repeat

    def __getitem__(self, data_id: int) -> Optional[Dict[str, Any]]:
        rec = self.samples[data_id % len(self.samples)]
        frames = self._load_video(rec["path"], only_used=True)
        keys = self.extract_middle_key_frames(frames)
        stitched = self.stitch_middle_key_frames(keys)
        if not frames:
            warnings.warn(f"cannot load frames from {rec['path']}")
            return None
        return {"image": frames[-1], "edit_image": frames[0], "middle_key_frames": keys, "stitched_image": stitched, "prompt": rec["prompt"],
                "state": rec["state"], "transition": rec["transition"], "idx": rec["idx"], "path": rec["path"], "original_prompt": rec["original_prompt"],
                "triplet": rec["triplet"], "supported_rules": rec["supported_rules"], "contradicted_rules": rec["contradicted_rules"]}


class Pica100kDataset(torch.utils.data.Dataset):
    """trainers/utils.py:685-775: image pairs of the PICA-100K set (Hugging Face `datasets` record: `src_img`, `tgt_img`, `superficial_prompt`) as
    training samples -- the target as `image`, the source as a one-element `edit_image` list, both cover-resized / centre-cropped like the clips."""

    def __init__(self, dataset_id: str = "Andrew613/PICA-100K", split: str = "train", cache_dir: Optional[str] = None, max_pixels: int = 1920 * 1080,
                 height: Optional[int] = None, width: Optional[int] = None, height_division_factor: int = 16, width_division_factor: int = 16,
                 repeat: int = 1, args=None):
        if args is not None:
            dataset_id = getattr(args, "dataset_id", dataset_id)
            height, width = getattr(args, "height", height), getattr(args, "width", width)
            max_pixels = getattr(args, "max_pixels", max_pixels)
            repeat = getattr(args, "dataset_repeat", repeat)
        self.dataset_id, self.split, self.cache_dir = dataset_id, split, cache_dir
        self.max_pixels, self.height, self.width, self.repeat = int(max_pixels), height, width, int(repeat)
        self.height_division_factor, self.width_division_factor = int(height_division_factor), int(width_division_factor)
        self.dynamic_resolution = resolution_mode(self.height, self.width)
        from datasets import load_dataset                     # local cache only on a machine without network: the caller's `cache_dir`
        self.data = load_dataset(self.dataset_id, split=self.split, cache_dir=self.cache_dir)

    _crop_and_resize = staticmethod(cover_and_center_crop)

    def _get_height_width(self, image: Image.Image) -> Tuple[int, int]:
        if not self.dynamic_resolution:
            return self.height, self.width
        return snapped_size(image, self.max_pixels, self.height_division_factor, self.width_division_factor)

    def _process_image(self, image: Image.Image) -> Image.Image:
        image = image.convert("RGB")
        return cover_and_center_crop(image, *self._get_height_width(image))

    def __len__(self) -> int:
        return len(self.data) * self.repeat

    def __getitem__(self, data_id: int) -> Optional[Dict[str, Any]]:
        rec = self.data[data_id % len(self.data)]
        src, tgt = rec.get("src_img"), rec.get("tgt_img")
        if src is None or tgt is None:
            warnings.warn("Pica100kDataset: missing src_img/tgt_img.")
            return None
        return {"image": self._process_image(tgt), "edit_image": [self._process_image(src)], "prompt": rec.get("superficial_prompt", "")}
