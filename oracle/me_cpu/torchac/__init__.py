"""CPU stand-in for torchac 0.9.3's two entry points.  TEST INFRASTRUCTURE ONLY (see oracle/rc_oracle.c)."""
import torch

from oracle import rc


def encode_float_cdf(cdf_float, sym, needs_normalization=True, check_input_bounds=False):
    return rc.encode_float_cdf(cdf_float.detach().cpu().numpy(), sym.detach().cpu().numpy())


def decode_float_cdf(cdf_float, byte_stream, needs_normalization=True):
    return torch.from_numpy(rc.decode_float_cdf(cdf_float.detach().cpu().numpy(), byte_stream))
