"""pyfilterbank_b200 -- B200-native (sm_100a) implementation of pyfilterbank's cascaded
second-order-section IIR filtering hot path, behind the reference's Python API."""
from .sosfiltering import (sosfilter_c, sosfilter_double_c, sosfilter_double_mimo_c,  # noqa: F401
                           bank_filter, bank_levels, bank_filter_weighted, bilinear_sos, get_bank)
from ._lib import Bank, PfbError  # noqa: F401
from .octbank import FractionalOctaveFilterbank  # noqa: F401,E402
from .gammatone import GammatoneFilterbank  # noqa: F401,E402
from .streaming import BlockSession  # noqa: F401,E402
from . import butterworth, octbank, sosfiltering, sharding, gammatone, splweighting, streaming  # noqa: F401,E402
