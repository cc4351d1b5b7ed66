"""`diffsynth.trainers.unified_dataset` (DiffSynth-Studio/diffsynth/trainers/unified_dataset.py): the generic metadata-driven dataset and its
composable loaders.  scripts/train/train_physicedit.py imports `UnifiedDataset` (:6); its role on this path is the READER of what
`launch_data_process_task` (`--task data_process`) caches: `UnifiedDataset(base_path=<cache dir>)` without a metadata file walks the directory
for `*.pth` files and yields the cached unit outputs, which `launch_training_task` feeds to the module as `model({}, inputs=data)`.

Host-side only.  Same class names, constructor arguments and `>>` composition as the reference, so operator pipelines written for it keep
working; images / frames are PIL objects, videos are decoded through `datasets.open_video` (imageio when installed, else OpenCV)."""
from __future__ import annotations

import json
import os

import torch
from PIL import Image

from .datasets import cover_and_center_crop, open_video


class DataProcessingOperator:
    """A callable step; `a >> b` is the step "a, then b" (:7-31)."""

    def __call__(self, data):
        raise NotImplementedError("DataProcessingOperator cannot be called directly.")

    def __rshift__(self, other):
        return DataProcessingPipeline(_steps(self) + _steps(other))


class DataProcessingPipeline(DataProcessingOperator):
    def __init__(self, operators=None):
        self.operators = [] if operators is None else operators

    def __call__(self, data):
        for op in self.operators:
            data = op(data)
        return data


def _steps(x):
    return list(x.operators) if isinstance(x, DataProcessingPipeline) else [x]


def _stateless(name: str, fn, doc: str):
    """An operator class that applies `fn` to its input (the reference spells each of these out as a class, :34-49, :111-114)."""
    return type(name, (DataProcessingOperator,), {"__call__": lambda self, data: fn(data), "__doc__": doc, "__module__": __name__, "__qualname__": name})


DataProcessingOperatorRaw = _stateless("DataProcessingOperatorRaw", lambda data: data, "identity")
ToInt = _stateless("ToInt", int, "int(data)")
ToFloat = _stateless("ToFloat", float, "float(data)")
ToList = _stateless("ToList", lambda data: [data], "[data]")


class ToStr(DataProcessingOperator):
    def __init__(self, none_value=""):
        self.none_value = none_value

    def __call__(self, data):
        return str(self.none_value if data is None else data)


class ToAbsolutePath(DataProcessingOperator):
    def __init__(self, base_path=""):
        self.base_path = base_path

    def __call__(self, data):
        return os.path.join(self.base_path, data)


class LoadImage(DataProcessingOperator):
    def __init__(self, convert_RGB=True):
        self.convert_RGB = convert_RGB

    def __call__(self, data: str):
        image = Image.open(data)
        return image.convert("RGB") if self.convert_RGB else image


class LoadTorchPickle(DataProcessingOperator):
    def __init__(self, map_location="cpu"):
        self.map_location = map_location

    def __call__(self, data):
        return torch.load(data, map_location=self.map_location, weights_only=False)


class ImageCropAndResize(DataProcessingOperator):
    """Fixed (height, width), or the image's own size scaled down to `max_pixels` and floored to the division factors; cover-resize + centre crop (:73-108)."""

    def __init__(self, height, width, max_pixels, height_division_factor, width_division_factor):
        self.height, self.width, self.max_pixels = height, width, max_pixels
        self.height_division_factor, self.width_division_factor = height_division_factor, width_division_factor

    crop_and_resize = staticmethod(cover_and_center_crop)

    def get_height_width(self, image):
        if self.height is not None and self.width is not None:
            return self.height, self.width
        width, height = image.size
        if width * height > self.max_pixels:
            scale = (width * height / self.max_pixels) ** 0.5
            height, width = int(height / scale), int(width / scale)
        return height // self.height_division_factor * self.height_division_factor, width // self.width_division_factor * self.width_division_factor

    def __call__(self, data: Image.Image):
        return cover_and_center_crop(data, *self.get_height_width(data))


class SequencialProcess(DataProcessingOperator):          # (sic) the reference's spelling is the interface
    def __init__(self, operator=lambda x: x):
        self.operator = operator

    def __call__(self, data):
        return [self.operator(item) for item in data]


def _frame_budget(total: int, num_frames: int, factor: int, remainder: int) -> int:
    """`num_frames`, or for a shorter clip the largest n <= its length with n % factor == remainder (:125-131, :164-171; no lower bound of one here)."""
    if total >= num_frames:
        return num_frames
    n = total
    while n > 1 and n % factor != remainder:
        n -= 1
    return n


class LoadVideo(DataProcessingOperator):
    def __init__(self, num_frames=81, time_division_factor=4, time_division_remainder=1, frame_processor=lambda x: x):
        self.num_frames, self.time_division_factor, self.time_division_remainder = num_frames, time_division_factor, time_division_remainder
        self.frame_processor = frame_processor                    # applied while decoding, frame by frame

    def get_num_frames(self, source):
        count = source.count_frames if hasattr(source, "count_frames") else source.count
        return _frame_budget(int(count()), self.num_frames, self.time_division_factor, self.time_division_remainder)

    def __call__(self, data: str):
        source = open_video(data)
        try:
            return [self.frame_processor(Image.fromarray(source.frame(i))) for i in range(self.get_num_frames(source))]
        finally:
            source.close()


class LoadGIF(DataProcessingOperator):
    def __init__(self, num_frames=81, time_division_factor=4, time_division_remainder=1, frame_processor=lambda x: x):
        self.num_frames, self.time_division_factor, self.time_division_remainder = num_frames, time_division_factor, time_division_remainder
        self.frame_processor = frame_processor

    @staticmethod
    def _frames(path):
        from PIL import ImageSequence
        with Image.open(path) as gif:
            return [frame.convert("RGB") for frame in ImageSequence.Iterator(gif)]

    def get_num_frames(self, path):
        return _frame_budget(len(self._frames(path)), self.num_frames, self.time_division_factor, self.time_division_remainder)

    def __call__(self, data: str):
        frames = self._frames(data)
        n = _frame_budget(len(frames), self.num_frames, self.time_division_factor, self.time_division_remainder)
        return [self.frame_processor(f) for f in frames[:max(n, 1)]]


class RouteByExtensionName(DataProcessingOperator):
    def __init__(self, operator_map):
        self.operator_map = operator_map

    def __call__(self, data: str):
        ext = data.split(".")[-1].lower()
        for names, op in self.operator_map:
            if names is None or ext in names:
                return op(data)
        raise ValueError(f"Unsupported file: {data}")


class RouteByType(DataProcessingOperator):
    def __init__(self, operator_map):
        self.operator_map = operator_map

    def __call__(self, data):
        for kind, op in self.operator_map:
            if kind is None or isinstance(data, kind):
                return op(data)
        raise ValueError(f"Unsupported data: {data}")


class UnifiedDataset(torch.utils.data.Dataset):
    """(:230-337) Two modes.  With `metadata_path` (.json list, .jsonl, or a CSV): one record per row, the fields named in `data_file_keys` are passed
    through `special_operator_map[key]` or `main_data_operator`.  Without it: `load_from_cache` -- every `*.pth` below `base_path`, loaded with torch."""

    def __init__(self, base_path=None, metadata_path=None, repeat=1, data_file_keys=tuple(), main_data_operator=lambda x: x, special_operator_map=None):
        self.base_path, self.metadata_path, self.repeat, self.data_file_keys = base_path, metadata_path, repeat, data_file_keys
        self.main_data_operator = main_data_operator
        self.cached_data_operator = LoadTorchPickle()
        self.special_operator_map = {} if special_operator_map is None else special_operator_map
        self.data, self.cached_data = [], []
        self.load_from_cache = metadata_path is None
        self.load_metadata(metadata_path)

    @staticmethod
    def default_image_operator(base_path="", max_pixels=1920 * 1080, height=None, width=None, height_division_factor=16, width_division_factor=16):
        one = ToAbsolutePath(base_path) >> LoadImage() >> ImageCropAndResize(height, width, max_pixels, height_division_factor, width_division_factor)
        return RouteByType(operator_map=[(str, one), (list, SequencialProcess(one))])

    @staticmethod
    def default_video_operator(base_path="", max_pixels=1920 * 1080, height=None, width=None, height_division_factor=16, width_division_factor=16,
                               num_frames=81, time_division_factor=4, time_division_remainder=1):
        fit = lambda: ImageCropAndResize(height, width, max_pixels, height_division_factor, width_division_factor)
        by_ext = RouteByExtensionName(operator_map=[
            (("jpg", "jpeg", "png", "webp"), LoadImage() >> fit() >> ToList()),
            (("gif",), LoadGIF(num_frames, time_division_factor, time_division_remainder, frame_processor=fit())),
            (("mp4", "avi", "mov", "wmv", "mkv", "flv", "webm"), LoadVideo(num_frames, time_division_factor, time_division_remainder, frame_processor=fit()))])
        return RouteByType(operator_map=[(str, ToAbsolutePath(base_path) >> by_ext)])

    def search_for_cached_data_files(self, path):
        for name in os.listdir(path):
            sub = os.path.join(path, name)
            if os.path.isdir(sub):
                self.search_for_cached_data_files(sub)
            elif sub.endswith(".pth"):
                self.cached_data.append(sub)

    def load_metadata(self, metadata_path):
        if metadata_path is None:
            print("No metadata_path. Searching for cached data files.")
            self.search_for_cached_data_files(self.base_path)
            print(f"{len(self.cached_data)} cached data files found.")
        elif metadata_path.endswith(".json"):
            with open(metadata_path, "r") as f:
                self.data = json.load(f)
        elif metadata_path.endswith(".jsonl"):
            with open(metadata_path, "r") as f:
                self.data = [json.loads(line.strip()) for line in f]
        else:
            import pandas
            table = pandas.read_csv(metadata_path)
            self.data = [table.iloc[i].to_dict() for i in range(len(table))]

    def __getitem__(self, data_id):
        if self.load_from_cache:
            return self.cached_data_operator(self.cached_data[data_id % len(self.cached_data)])
        data = self.data[data_id % len(self.data)].copy()
        for key in self.data_file_keys:
            if key in data:
                data[key] = self.special_operator_map.get(key, self.main_data_operator)(data[key])
        return data

    def __len__(self):
        return (len(self.cached_data) if self.load_from_cache else len(self.data)) * self.repeat

    def check_data_equal(self, data1, data2):
        return len(data1) == len(data2) and all(data1[k] == data2[k] for k in data1)
