"""`utils.tfrecord_voc_utils` of the reference, TensorFlow-free (ref utils/tfrecord_voc_utils.py).

`get_generator(tfrecords, batch_size, buffer_size, image_preprocess_config)` returns the reference's
`(init_op, iterator)` pair (:115-120): `init_op()` (re)starts the stream, `iterator.get_next()` yields
`(images float32 [B,H,W,3] or [B,3,H,W], ground_truth float32 [B,pad_truth_to,5])` batches read from the
TFRecord files with odt_b200.tfrecord (TFRecord framing, tf.train.Example, OpenCV JPEG decoding, the
deterministic resize path of image_augmentor).  Nothing is opened until the first `get_next()`, so the
drivers' construction order (`get_generator` before the model, testSSD300.py:48-60) works even when the
data directory is empty.  `dataset2tfrecord` needs the VOC XML parser of the reference and stays out.
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
