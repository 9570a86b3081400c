// qboot_b200 façade -- the free functions of include/qboot/hor_recursion.hpp:19-30.  gBlock is declared with Context
// (qboot/context.hpp): one table, computed on the GPU.  The intermediate stages of the reference (hBlock_shifted,
// gBlock_real) are kernels of the engine (k_series, k_scale / k_horner, k_rho_to_z) and have no host entry points.
#ifndef QBOOT_B200_HOR_RECURSION_HPP_
#define QBOOT_B200_HOR_RECURSION_HPP_
#include "qboot/context.hpp"
#endif  // QBOOT_B200_HOR_RECURSION_HPP_
