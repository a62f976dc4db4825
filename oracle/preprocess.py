"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the collate-time histogram standardisation,
classification/train_ENC_CLF.ipynb [cell 9] (`_standardize_cutoff`, `_get_percentiles`, `normalize`) in numpy float64, and of
`reshape_image` (utils/data.py:16-30).  Pinned against the notebook's own functions by tests/golden/histstd_*.npz
(oracle/make_golden.py `histstd_cases`; the notebook cell is exec'd with `np.bool = bool`, which numpy >= 1.24 removed)."""
import numpy as np

DEFAULT_CUTOFF = 0.01, 0.99
RANGE_TO_USE = [0, 1, 2, 4, 5, 6, 7, 8, 10, 11, 12]


def standardize_cutoff(cutoff):                     # [cell 9] _standardize_cutoff
    cutoff = np.asarray(cutoff, dtype=np.float64).copy()
    cutoff[0] = max(0., cutoff[0])
    cutoff[1] = min(1., cutoff[1])
    cutoff[0] = np.min([cutoff[0], 0.09])
    cutoff[1] = np.max([cutoff[1], 0.91])
    return cutoff


def get_percentiles(percentiles_cutoff):            # [cell 9] _get_percentiles
    quartiles = np.arange(25, 100, 25).tolist()
    deciles = np.arange(10, 100, 10).tolist()
    return np.array(sorted(set(list(percentiles_cutoff) + quartiles + deciles)))


def percentile_values(array, mask=None, cutoff=None):
    data = np.asarray(array).reshape(-1).astype(np.float32)
    m = np.ones_like(data, bool) if mask is None else np.asarray(mask).reshape(-1).astype(bool)
    percentiles = get_percentiles(100 * np.array(standardize_cutoff(DEFAULT_CUTOFF if cutoff is None else cutoff)))
    return np.percentile(data[m], percentiles)


def normalize(array, landmarks, mask=None, cutoff=None, epsilon=1e-5):      # [cell 9] normalize
    array = np.asarray(array)
    shape = array.shape
    data = array.reshape(-1).astype(np.float32)
    pv = percentile_values(array, mask, cutoff)
    mapping = np.asarray(landmarks)
    range_mapping = mapping[RANGE_TO_USE]
    range_perc = pv[RANGE_TO_USE]
    diff_mapping = np.diff(range_mapping)
    diff_perc = np.diff(range_perc)
    diff_perc[diff_perc < epsilon] = np.inf
    affine_map = np.zeros([2, len(RANGE_TO_USE) - 1])
    affine_map[0] = diff_mapping / diff_perc
    affine_map[1] = range_mapping[:-1] - affine_map[0] * range_perc[:-1]
    bin_id = np.digitize(data, range_perc[1:-1], right=False)
    new_img = affine_map[0, bin_id] * data + affine_map[1, bin_id]
    return new_img.reshape(shape).astype(np.float32)


def reshape_image(img, coord_min, img_shape):       # utils/data.py:16-30
    img_shape = tuple(img_shape)
    img = img[coord_min[0]:coord_min[0] + img_shape[0], coord_min[1]:coord_min[1] + img_shape[1], coord_min[2]:coord_min[2] + img_shape[2]]
    if img.shape[:3] != img_shape:
        raise AssertionError
    return img.reshape((1,) + img_shape)
