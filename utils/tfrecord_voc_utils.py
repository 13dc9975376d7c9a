"""Import-time stand-in for the reference's TFRecord input pipeline
(ref utils/tfrecord_voc_utils.py:115-120).  Every reference driver calls
`voc_utils.get_generator(...)` before constructing the model and hands the
returned 2-tuple to the constructor as data_provider['train_generator']; the
inference hot path never consumes it.  The real TFRecord/JPEG/augmentation
pipeline is training input and out of scope here.
"""


class _Generator:
    def __init__(self, tfrecords, batch_size, buffer_size, config):
        self.tfrecords, self.batch_size = tfrecords, batch_size
        self.buffer_size, self.config = buffer_size, config

    def get_next(self):
        raise NotImplementedError("the TFRecord training input pipeline is outside the "
                                  "accelerated inference hot path (SURVEY.md 8f)")


def get_generator(tfrecords, batch_size, buffer_size, image_preprocess_config):
    """Returns (init_op, iterator) like the reference; both are inert placeholders."""
    it = _Generator(tfrecords, batch_size, buffer_size, image_preprocess_config)
    return None, it
