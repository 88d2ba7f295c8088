// pfb_inst_fd.cu -- kernels for signal type float, arithmetic type double (see pfb_inst.inl).
#define PFB_T float
#define PFB_A double
#define PFB_SUFFIX fd
#define PFB_HAS_EXACT false   // the exact (reference rounding) variant only exists when T == A
#define PFB_HAS_SCALED 1   // gain-normalised TMA kernels need float64 arithmetic
#define PFB_HAS_DEC 0      // decimating-band output pass (the decimating band path is float64)
#include "pfb_inst.inl"
