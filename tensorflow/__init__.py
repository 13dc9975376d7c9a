"""Import shim: the reference drivers do `import tensorflow as tf` but never use
`tf.` afterwards.  There is no TensorFlow runtime on this
path; touching any attribute is an error."""


def __getattr__(name):
    raise AttributeError("tensorflow shim: the B200 path has no TensorFlow runtime (asked for tf.%s)"
                         % name)
