"""Stand-in for scikit-image 0.21.0 (environment.yaml:77; NOT importable in this image) so that the reference's
model_v1/data/transform.py and dataset.py can be exec'd unmodified by oracle/make_augment_golden.py.
TEST INFRASTRUCTURE ONLY.  Only the symbols those two files touch exist; their arithmetic is the restatement in
oracle/augment_oracle.py (parity unpinned, see its header)."""
import numpy as np

from . import transform  # noqa: F401


def img_as_float32(image):                      # dataset.py:125: uint8 -> float32 in [0, 1]
    image = np.asarray(image)
    if image.dtype == np.uint8:
        return image.astype(np.float32) / np.float32(255)
    return image.astype(np.float32)
