// Kernel variant 16 of libtactilesim_b200.so: <= 16 reduced dofs, 16 lanes per environment, cuboid and cylinder
// primitives (DClaw, TactileInsertion).  Capacities: kernel_layout.h.
#define TS_VARIANT 16
#include "kernels.cu"
