// Kernel variant 17 of libtactilesim_b200.so: the capacities of variant 16 plus sphere primitives, free3d-exp
// joints, BDF2 / SDIRK2 integration (forward only; the adjoint is BDF1's) and dense contact point sets (the rolling-ball
// scene of examples/RollingBallExp, BASELINE configs[0]).  Capacities: kernel_layout.h.
#define TS_VARIANT 17
#include "kernels.cu"
