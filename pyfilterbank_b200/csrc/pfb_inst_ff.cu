// pfb_inst_ff.cu -- kernels for signal type float, arithmetic type float (see pfb_inst.inl).
#define PFB_T float
#define PFB_A float
#define PFB_SUFFIX ff
#define PFB_HAS_EXACT true   // the exact (reference rounding) variant only exists when T == A
#define PFB_HAS_SCALED 0   // gain-normalised TMA kernels need float64 arithmetic
#define PFB_HAS_DEC 0      // decimating-band output pass (the decimating band path is float64)
#include "pfb_inst.inl"
