from .augment import LetterBox  # noqa: F401
from .loaders import LoadImagesAndVideos, decode_jpeg, decode_jpeg_batch, imread_device  # noqa: F401
from .dataset import YOLODataset, build_yolo_dataset, check_det_dataset, img2label_paths  # noqa: F401
