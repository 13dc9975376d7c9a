"""`utils.tfrecord_voc_utils` of the reference, TensorFlow-free (ref utils/tfrecord_voc_utils.py).

`get_generator(tfrecords, batch_size, buffer_size, image_preprocess_config)` returns the reference's
`(init_op, iterator)` pair (:115-120): `init_op()` (re)starts the stream, `iterator.get_next()` yields
`(images float32 [B,H,W,3] or [B,3,H,W], ground_truth float32 [B,pad_truth_to,5])` batches read from the
TFRecord files with odt_b200.tfrecord (TFRecord framing, tf.train.Example, OpenCV JPEG decoding,
image_augmentor's resize / zoom / crop / flip / box arithmetic).  Nothing is opened until the first `get_next()`, so the
drivers' construction order (`get_generator` before the model, testSSD300.py:48-60) works even when the
data directory is empty.  `dataset2tfrecord` / `xml_to_example` (:30-92) convert a VOC annotation
directory into the same sharded records with the standard-library XML parser.
"""
import os
import sys

_PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "object-detection-tensorflow_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)


def get_generator(tfrecords, batch_size, buffer_size, image_preprocess_config):
    from odt_b200.tfrecord import BatchIterator
    it = BatchIterator(tfrecords, batch_size, buffer_size, image_preprocess_config)
    return it.initialize, it


def xml_to_example(xmlpath, imgpath):
    """One VOC annotation -> serialized tf.train.Example (ref :30-62): 'image' = the JPEG file's bytes,
    'shape' = int32 (height, width, depth), 'ground_truth' = float32 [n,5] rows (ymin, ymax, xmin, xmax, id)."""
    import xml.etree.ElementTree as ET

    import numpy as np
    from odt_b200.tfrecord import encode_voc_example
    from utils.voc_classname_encoder import classname_to_ids
    root = ET.parse(xmlpath).getroot()
    with open(os.path.join(imgpath, root.find("filename").text), "rb") as f:
        image = f.read()
    size = root.find("size")
    shape = [int(size.find(k).text) for k in ("height", "width", "depth")]
    objs = root.findall(".//object")
    gt = np.zeros((len(objs), 5), np.float32)
    for i, obj in enumerate(objs):
        bb = obj.find("bndbox")
        gt[i] = [float(bb.find("ymin").text), float(bb.find("ymax").text), float(bb.find("xmin").text),
                 float(bb.find("xmax").text), classname_to_ids[obj.find("name").text]]
    return encode_voc_example(image, shape, gt)


def dataset2tfrecord(xml_dir, img_dir, output_dir, name, total_shards=5):
    """ref :65-92: `<name>_%05d-of-%05d.tfrecord` shards of the annotations in `xml_dir` (the reference's
    shard arithmetic -- ceil applied to the integer count before the division -- is kept)."""
    import glob
    import math

    from odt_b200.tfrecord import write_records
    os.makedirs(output_dir, exist_ok=True)
    xmllist = sorted(glob.glob(os.path.join(xml_dir, "*.xml")))
    num_per_shard = int(math.ceil(len(xmllist)) / float(total_shards))
    outputfiles = []
    for shard_id in range(total_shards):
        out = os.path.join(output_dir, "%s_%05d-of-%05d.tfrecord" % (name, shard_id + 1, total_shards))
        outputfiles.append(out)
        lo, hi = shard_id * num_per_shard, min((shard_id + 1) * num_per_shard, len(xmllist))
        write_records(out, (xml_to_example(xmllist[i], img_dir) for i in range(lo, hi)))
    return outputfiles
