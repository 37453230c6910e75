// Kernel variant 8 of libtactilesim_b200.so: <= 8 reduced dofs, 8 lanes per environment, cuboid primitives
// (TactilePush).  Capacities: kernel_layout.h.  A separate file name keeps the two cubins apart in the fatbinary.
#define TS_VARIANT 8
#include "kernels.cu"
