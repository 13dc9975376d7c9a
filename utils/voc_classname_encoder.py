"""PASCAL-VOC class name <-> id map (ids are the contract of `class_id` in
test_one_image results; ref utils/voc_classname_encoder.py:1-22)."""
_NAMES = ["aeroplane", "bicycle", "bird", "boat", "bottle", "bus", "car", "cat", "chair", "cow",
          "diningtable", "dog", "horse", "motorbike", "person", "pottedplant", "sheep", "sofa",
          "train", "tvmonitor"]
classname_to_ids = {n: i for i, n in enumerate(_NAMES)}
