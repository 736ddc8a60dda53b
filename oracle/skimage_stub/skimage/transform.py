"""skimage.transform symbols used by the reference's RandomTransform (model_v1/data/transform.py:164-230):
ProjectiveTransform (estimate / inverse / __add__ / params), SimilarityTransform(translation=...), warp, resize.
Restated from scikit-image 0.21.0 in oracle/augment_oracle.py.  TEST INFRASTRUCTURE ONLY."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import augment_oracle as A  # noqa: E402


class ProjectiveTransform(object):
    def __init__(self, matrix=None):
        self.params = np.eye(3) if matrix is None else np.asarray(matrix, dtype=np.float64)

    def estimate(self, src, dst):
        self.params = A.homography(src, dst)
        return True

    def __call__(self, coords):
        return A.apply_h(self.params, coords)

    @property
    def inverse(self):
        return ProjectiveTransform(np.linalg.inv(self.params))

    def __add__(self, other):
        return ProjectiveTransform(other.params @ self.params)


class SimilarityTransform(ProjectiveTransform):
    def __init__(self, translation=(0, 0)):
        tx, ty = translation
        super().__init__(np.array([[1, 0, tx], [0, 1, ty], [0, 0, 1]], dtype=np.float64))


def warp(image, inverse_map, output_shape=None, cval=0.0, preserve_range=False):
    assert preserve_range
    shp = tuple(int(v) for v in np.asarray(output_shape))
    return A.warp_projective(np.asarray(image), inverse_map.params, shp, cval)


def resize(image, output_shape, preserve_range=False):
    assert preserve_range
    return A.resize_like_skimage(np.asarray(image), output_shape)
