from .augment import LetterBox  # noqa: F401
from .loaders import LoadImagesAndVideos, decode_jpeg, imread_device  # noqa: F401
