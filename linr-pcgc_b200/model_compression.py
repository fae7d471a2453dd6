"""8-bit model quantisation and its bitstream: host-side mirror of `Model_Estimate`
(model_compression/model_size_est.py:72-91 quant_uniform2, :390-519 compress_model, :523-579 decompress_model).

Quantisation, min/max and the Laplace statistics are one kernel over the flat parameter vector
(`linr_param_quant`); mode selection, zlib and the range coder stay on the host as in the reference.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict

import numpy as np
import torch

from . import rc
from .net import param_quant


CDF_REFERENCE, CDF_FIXED = 1, 2   # `cdf_version` of the Laplace model coder (side info / container header)


def laplace_cdf_row(mu: float, b: float, bitdepth: int = 8, cdf_version: int = CDF_REFERENCE) -> np.ndarray:
    """16-bit CDF row of the Laplace model coder.
    version 1 (the reference's bitstream): float row `[c1..c_L, 0]` (model_size_est.py:468-478 -- the cumulative sum has no
    leading zero and a trailing zero, so symbol s is coded with the mass of s+1; kept as is for bitstream parity).
    version 2: the intended row `[0, c1..c_L]` (symbol s coded with its own mass); same coder, smaller model.bin."""
    L = int(math.ceil(2 ** bitdepth))
    x = torch.arange(L, dtype=torch.float32)
    pdf = torch.exp(-torch.abs(x - mu) / b) / (2 * b)
    pdf = pdf / pdf.sum()
    cdf = torch.cumsum(pdf, dim=-1).to(torch.float32)
    if cdf_version == CDF_REFERENCE:
        row = torch.cat([cdf, torch.zeros(1)]).numpy()
    elif cdf_version == CDF_FIXED:
        row = torch.cat([torch.zeros(1), cdf]).clamp(max=1.0).numpy()
    else:
        raise ValueError(f"unknown cdf_version {cdf_version}")
    return rc.cdf_float_to_u16(row)


def _sym_dtype(bitdepth: int):
    return np.uint8 if bitdepth <= 8 else np.dtype("<u2")


def compress_model(flat: torch.Tensor, bitdepth: int = 8, cdf_version: int = CDF_REFERENCE) -> Dict:
    """flat: CUDA fp32 parameter vector -> dict with the reference's keys (`enc_mode, final_bytes, bit_real, mu, b,
    min_param, max_param, recon_ret, bitdepth`).  `recon_ret` (CUDA) is the dequantised model the encoder codes with
    (encoder.py:101-103).  Bit depths 9..16 use 16-bit symbols in the raw / zlib modes (the reference skips the range
    coder above 8 bits, model_size_est.py:463,490-491, and cannot decode what it wrote, :546-548)."""
    if not 1 <= bitdepth <= 16:
        raise ValueError("model_bitdepth must be in [1,16]")
    q, recon, stats = param_quant(flat, bitdepth)
    mn, mx, mu, b = [float(v) for v in stats.cpu().tolist()]
    n = int(q.numel())
    qf = q.to(torch.float32) if bitdepth <= 8 else (q.to(torch.int32) & 0xFFFF).to(torch.float32)
    like = torch.exp(-torch.abs(qf - mu) / b) / (2 * b)           # mylaplace_pdf on the symbols
    bits = float((-torch.sum(torch.log2(like))).item()) + 2 * bitdepth
    bpp = bits / n
    q_u8 = q.cpu().numpy() if bitdepth <= 8 else q.cpu().numpy().view(np.uint16).astype("<u2")   # "u8": the symbol array
    q_bytes = q_u8.tobytes()
    q_zlib = zlib.compress(q_bytes)
    bpp_zlib = len(q_zlib) * 8 / n
    low_bound = bpp_zlib if bpp_zlib < bitdepth else bitdepth

    def fallback():
        if low_bound == bitdepth:
            return 0, q_bytes
        return 1, q_zlib

    bit_lap = float("inf")
    if bpp > low_bound or bitdepth > 8:
        enc_mode, final = fallback()
        bit_real = low_bound * n + 2 + 64
    else:
        ac = rc.encode_shared(laplace_cdf_row(mu, b, bitdepth, cdf_version), q_u8.astype(np.int16))
        bit_lap = len(ac) * 8 + 2 * math.ceil(bitdepth) + 2 + 64
        if bit_lap > low_bound * n + 2 + 64:
            enc_mode, final = fallback()
            bit_real = low_bound * n + 2 + 64
        else:
            enc_mode, final, bit_real = 2, ac, bit_lap
    return dict(enc_mode=enc_mode, final_bytes=final, bit_real=bit_real, bpp_real=bit_real / n, mu=mu, b=b, min_param=mn,
                max_param=mx, recon_ret=recon, quant=q, bitdepth=bitdepth, zlib_bpp=bpp_zlib, laplace_real_bpp=bit_lap / n,
                cdf_version=cdf_version)


def decompress_model(enc: Dict, n: int, device="cuda") -> torch.Tensor:
    """Inverse: side info (`mu, b, min_param, max_param, enc_mode, bitdepth`) + bytes -> flat fp32 vector on `device`,
    bit-identical to `compress_model(...)['recon_ret']`."""
    bd = int(enc["bitdepth"])
    mode = int(enc["enc_mode"])
    data = enc["final_bytes"]
    if mode == 0:
        q = np.frombuffer(data, dtype=_sym_dtype(bd))
    elif mode == 1:
        q = np.frombuffer(zlib.decompress(data), dtype=_sym_dtype(bd))
    else:
        q = rc.decode_shared(laplace_cdf_row(float(enc["mu"]), float(enc["b"]), bd, int(enc.get("cdf_version", CDF_REFERENCE))), data, n)
    if len(q) != n:
        raise ValueError(f"model bitstream holds {len(q)} symbols, model has {n} parameters")
    smax = np.float32(math.ceil(2 ** bd) - 1)
    mn, mx = np.float32(enc["min_param"]), np.float32(enc["max_param"])
    recon = q.astype(np.float32) / smax * (mx - mn) + mn    # fp32, same op order as quant_uniform2 (model_size_est.py:88)
    return torch.from_numpy(recon.astype(np.float32)).to(device)
