// qboot_b200 façade -- include/qboot/hor_formula.hpp declares _get_rec_coeffs, the recurrence coefficients of the HOR
// series.  On this engine they are produced on the device (k_rec_coeffs over qboot_b200/csrc/hor_table.inc); the header
// exists so that sources including it keep compiling.
#ifndef QBOOT_B200_HOR_FORMULA_HPP_
#define QBOOT_B200_HOR_FORMULA_HPP_
#include "qboot/primary_op.hpp"
#endif  // QBOOT_B200_HOR_FORMULA_HPP_
