from .object_detection import ObjectDetection  # noqa: F401
