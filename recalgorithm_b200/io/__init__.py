"""Wire formats on the feeding side of the hot path: TFRecord framing, tf.train.Example /
SequenceExample protobuf wire parsing, vocabulary files.  Host-side Python; no TensorFlow, no protobuf codegen."""
from .tfrecord import masked_crc32c, read_records, write_records  # noqa: F401
from .example import (encode_example, encode_sequence_example, parse_single, parse_example,  # noqa: F401
                      FixedLenFeature, VarLenFeature)
from .vocab import VocabularyFile  # noqa: F401
