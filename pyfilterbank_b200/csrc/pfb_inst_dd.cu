// pfb_inst_dd.cu -- kernels for signal type double, arithmetic type double (see pfb_inst.inl).
#define PFB_T double
#define PFB_A double
#define PFB_SUFFIX dd
#define PFB_HAS_EXACT true   // the exact (reference rounding) variant only exists when T == A
#define PFB_HAS_SCALED 1   // gain-normalised TMA kernels need float64 arithmetic
#define PFB_HAS_DEC 1      // decimating-band output pass (the decimating band path is float64)
#include "pfb_inst.inl"
