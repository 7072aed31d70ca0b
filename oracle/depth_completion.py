"""Oracle: classical multi-scale depth completion (the ip_basic algorithm).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows the reference
``models/utils/ip_basic/depth_map_utils.py:134-287`` (``fill_in_multiscale``) for
the only configuration the hot path uses (``extrapolate=False``,
``blur_type='bilateral'``, reference encoder_utils.py:175-182), with OpenCV doing
the morphology exactly as in the reference.  ``fill_in_multiscale_numpy`` is an
OpenCV-free restatement of the same filters (border rules per SURVEY.md App. C.7)
used to pin the CUDA kernel's semantics stage by stage.
"""
import numpy as np

try:  # OpenCV is what the reference calls; keep it optional for box images without it
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

MAX_DEPTH = 100.0


def _cross(n):
    k = np.zeros((n, n), np.uint8)
    k[n // 2, :] = 1
    k[:, n // 2] = 1
    return k


def _full(n):
    return np.ones((n, n), np.uint8)


def fill_in_multiscale(depth_map, return_stages=False):
    """depth_map (h,w) float -> completed (h,w) float32.  reference :134-287."""
    assert cv2 is not None, "oracle depth completion needs OpenCV"
    d0 = np.float32(depth_map)
    near = (d0 > 0.1) & (d0 <= 15.0)
    med = (d0 > 15.0) & (d0 <= 30.0)
    far = d0 > 30.0
    # :170-174  invert valid depths
    s1 = d0.copy()
    v = s1 > 0.1
    s1[v] = MAX_DEPTH - s1[v]
    # :176-196  per-bin dilation with cross kernels 3 / 5 / 7, merged far -> near
    dil_far = cv2.dilate(s1 * far, _cross(3))
    dil_med = cv2.dilate(s1 * med, _cross(5))
    dil_near = cv2.dilate(s1 * near, _cross(7))
    s2 = s1.copy()
    for dil in (dil_far, dil_med, dil_near):
        m = dil > 0.1
        s2[m] = dil[m]
    # :198-200  5x5 closing
    s3 = cv2.morphologyEx(s2, cv2.MORPH_CLOSE, _full(5))
    # :202-206  median where valid
    s4 = s3.copy()
    med5 = cv2.medianBlur(s3, 5)
    v = s3 > 0.1
    s4[v] = med5[v]
    # :208-222  fill empties below the top-most valid pixel of each column with a 9x9 dilation
    top = _top_mask(s4)
    empty = ~(s4 > 0.1) & top
    dil9 = cv2.dilate(s4, _full(9))
    s5 = s4.copy()
    s5[empty] = dil9[empty]
    # :224-238  (extrapolate=False) recompute the top mask on s5
    top = _top_mask(s5)
    # :240-245  six masked 5x5 dilations
    s7 = s5.copy()
    for _ in range(6):
        empty = (s7 < 0.1) & top
        dil5 = cv2.dilate(s7, _full(5))
        s7[empty] = dil5[empty]
    s6 = s7.copy()
    # :247-250  median where valid; NOTE the mask is taken BEFORE the median is applied
    med5 = cv2.medianBlur(s7, 5)
    v = (s7 > 0.1) & top
    s7[v] = med5[v]
    s7m = s7.copy()
    # :257-260  bilateral written at the same (pre-median) mask
    bil = cv2.bilateralFilter(s7, 5, 0.5, 2.0)
    s7[v] = bil[v]
    # :262-266  re-invert
    out = s7.copy()
    vv = out > 0.1
    out[vv] = MAX_DEPTH - out[vv]
    if return_stages:
        return out, dict(s1=s1, s2=s2, s3=s3, s4=s4, s5=s5, s6=s6, s7m=s7m, s7=s7)
    return out


def _top_mask(img):
    """True at and below the first row with value > 0.1 in each column
    (reference :209-213, :226-238; a column without valid pixels has argmax 0 =>
    all True)."""
    first = np.argmax(img > 0.1, axis=0)
    rows = np.arange(img.shape[0])[:, None]
    return rows >= first[None, :]


# ----------------------------------------------------------------------------------------------
# OpenCV-free restatement (same semantics, used to pin the CUDA kernel stage by stage)
# ----------------------------------------------------------------------------------------------

def _offsets(kernel):
    r = kernel.shape[0] // 2
    return [(i - r, j - r) for i in range(kernel.shape[0]) for j in range(kernel.shape[1]) if kernel[i, j]]


def _dilate(img, kernel):
    """cv2.dilate, default border: out-of-image taps never win (== -inf)."""
    h, w = img.shape
    r = kernel.shape[0] // 2
    pad = np.full((h + 2 * r, w + 2 * r), -np.inf, np.float32)
    pad[r:r + h, r:r + w] = img
    out = np.full((h, w), -np.inf, np.float32)
    for dy, dx in _offsets(kernel):
        out = np.maximum(out, pad[r + dy:r + dy + h, r + dx:r + dx + w])
    return out


def _erode(img, kernel):
    return -_dilate(-img, kernel)


def _median5(img):
    """cv2.medianBlur(float32, 5): replicate border, exact median of 25."""
    h, w = img.shape
    pad = np.pad(img, 2, mode='edge')
    stack = np.stack([pad[i:i + h, j:j + w] for i in range(5) for j in range(5)], 0)
    return np.sort(stack, axis=0)[12]


def bilateral5(img, sigma_color=0.5, sigma_space=2.0):
    """cv2.bilateralFilter(float32, d=5): reflect-101 border, circular support
    r<=2 (13 taps incl. centre), colour weight from a 4096-bin LUT over
    [0, max-min] with linear interpolation (OpenCV imgproc/bilateral_filter)."""
    h, w = img.shape
    radius = 2
    mn, mx = float(img.min()), float(img.max())
    if abs(mx - mn) < np.finfo(np.float32).eps:
        return img.copy()
    gauss_color = -0.5 / (sigma_color * sigma_color)
    gauss_space = -0.5 / (sigma_space * sigma_space)
    nbins = 1 << 12
    ln = np.float32(mx - mn)
    scale_index = np.float32(nbins / ln)
    lut = np.zeros(nbins + 2, np.float32)
    last = 1.0
    for i in range(nbins + 2):
        if last > 0.0:
            val = i / float(scale_index)
            lut[i] = np.float32(np.exp(val * val * gauss_color))
            last = lut[i]
    pad = np.pad(img, radius, mode='reflect')
    s = np.zeros((h, w), np.float32)
    ws = np.zeros((h, w), np.float32)
    for i in range(-radius, radius + 1):
        for j in range(-radius, radius + 1):
            r = np.sqrt(float(i * i + j * j))
            if r > radius:
                continue
            sw = np.float32(np.exp(r * r * gauss_space))
            val = pad[radius + i:radius + i + h, radius + j:radius + j + w]
            alpha = np.abs(val - img) * scale_index
            idx = np.floor(alpha).astype(np.int64)
            alpha = (alpha - idx).astype(np.float32)
            wgt = sw * (lut[idx] + alpha * (lut[idx + 1] - lut[idx]))
            s += val * wgt
            ws += wgt
    return (s / ws).astype(np.float32)


def fill_in_multiscale_numpy(depth_map, return_stages=False):
    """Same pipeline as :func:`fill_in_multiscale` without OpenCV."""
    d0 = np.float32(depth_map)
    near = (d0 > 0.1) & (d0 <= 15.0)
    med = (d0 > 15.0) & (d0 <= 30.0)
    far = d0 > 30.0
    s1 = d0.copy()
    v = s1 > 0.1
    s1[v] = MAX_DEPTH - s1[v]
    dil_far = _dilate(s1 * far, _cross(3))
    dil_med = _dilate(s1 * med, _cross(5))
    dil_near = _dilate(s1 * near, _cross(7))
    s2 = s1.copy()
    for dil in (dil_far, dil_med, dil_near):
        m = dil > 0.1
        s2[m] = dil[m]
    s3 = _erode(_dilate(s2, _full(5)), _full(5))
    s4 = s3.copy()
    m5 = _median5(s3)
    v = s3 > 0.1
    s4[v] = m5[v]
    top = _top_mask(s4)
    empty = ~(s4 > 0.1) & top
    d9 = _dilate(s4, _full(9))
    s5 = s4.copy()
    s5[empty] = d9[empty]
    top = _top_mask(s5)
    s7 = s5.copy()
    for _ in range(6):
        empty = (s7 < 0.1) & top
        d5 = _dilate(s7, _full(5))
        s7[empty] = d5[empty]
    s6 = s7.copy()
    m5 = _median5(s7)
    v = (s7 > 0.1) & top
    s7[v] = m5[v]
    s7m = s7.copy()
    bil = bilateral5(s7)
    s7[v] = bil[v]
    out = s7.copy()
    vv = out > 0.1
    out[vv] = MAX_DEPTH - out[vv]
    if return_stages:
        return out, dict(s1=s1, s2=s2, s3=s3, s4=s4, s5=s5, s6=s6, s7m=s7m, s7=s7)
    return out
