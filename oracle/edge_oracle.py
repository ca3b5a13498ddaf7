"""CPU restatement of the reference's WaveletsHFBlock (TEST INFRASTRUCTURE: only tests/, smoke() and bench.py's CPU legs may
import this; the product never does).

Reference: yolox/models/surrounding_extraction.py
  DWT_2D / DWT_Function.forward   :12-29,103-124   stride-2 depthwise cross-correlation with the four 2x2 Haar filters
                                                   w_xy[i][j] = dec_x_reversed[i] * dec_y_reversed[j] ... (rows = first index)
  IDWT_2D / IDWT_Function.forward :43-59,77-100    stride-2 transposed depthwise convolution with rec_* filters
  WaveletsHFBlock.forward         :257-267         LF zeroed, HF -> Conv1x1(3C,3C)+ReLU, inverse transform,
                                                   times Conv3x3(C,C,pad 1)+ReLU of the input
pinned by tests/golden/edge.npz (outputs of the reference module, tools/make_goldens_host.py)."""
import numpy as np
import torch
import torch.nn.functional as F

# pywt.Wavelet('haar'): 1/sqrt(2) taps; the reference builds its 2x2 filters as fp32 products of fp32 taps
_S = np.float32(1.0 / np.sqrt(2.0))
_K = float(_S * _S)


def haar_dwt_hf(x: torch.Tensor):
    """[B,C,H,W] -> (LH, HL, HH) each [B,C,H/2,W/2]; sub-band order of DWT_Function.forward's concat (:24-28)."""
    a, b = x[:, :, 0::2, 0::2], x[:, :, 0::2, 1::2]
    c, d = x[:, :, 1::2, 0::2], x[:, :, 1::2, 1::2]
    lh = _K * a + _K * b - _K * c - _K * d        # w_lh[i][j] = dec_hi_r[i] * dec_lo_r[j]: high-pass over rows
    hl = _K * a - _K * b + _K * c - _K * d        # w_hl[i][j] = dec_lo_r[i] * dec_hi_r[j]: high-pass over columns
    hh = _K * a - _K * b - _K * c + _K * d
    return lh, hl, hh


def haar_idwt(ll, lh, hl, hh):
    """Inverse transform (IDWT_Function.forward :49-59): out[2Y+i, 2X+j] = sum_s band_s[Y,X] * rec_filter_s[i][j]."""
    B, C, H, W = lh.shape
    out = lh.new_zeros(B, C, 2 * H, 2 * W)
    for i in (0, 1):
        for j in (0, 1):
            si, sj = (1.0 if i == 0 else -1.0), (1.0 if j == 0 else -1.0)
            out[:, :, i::2, j::2] = _K * ll + (_K * si) * lh + (_K * sj) * hl + (_K * si * sj) * hh
    return out


def wavelets_hf_block(x: torch.Tensor, w1, b1, w3, b3) -> torch.Tensor:
    """WaveletsHFBlock.forward (:257-267).  x [B,C,H,W] (H, W even); w1 [3C,3C,1,1], w3 [C,C,3,3]."""
    C = x.shape[1]
    lh, hl, hh = haar_dwt_hf(x)
    hf = F.relu(F.conv2d(torch.cat([lh, hl, hh], 1), w1, b1))           # filter1 on the HF channels [LH|HL|HH]
    lh2, hl2, hh2 = hf.split([C, C, C], 1)
    x_idwt = haar_idwt(torch.zeros_like(lh2), lh2, hl2, hh2)            # LF = 0 (:261)
    x_content = F.relu(F.conv2d(x, w3, b3, padding=1))                  # filter2
    return x_content * x_idwt
