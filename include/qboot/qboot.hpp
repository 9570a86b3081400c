// qboot_b200 façade -- umbrella header: `#include "qboot/qboot.hpp"` as with the reference (include/qboot/qboot.hpp).
#ifndef QBOOT_B200_QBOOT_HPP_
#define QBOOT_B200_QBOOT_HPP_
#include "qboot/algebra/complex_function.hpp"
#include "qboot/algebra/matrix.hpp"
#include "qboot/algebra/polynomial.hpp"
#include "qboot/algebra/real_function.hpp"
#include "qboot/block.hpp"
#include "qboot/bootstrap_equation.hpp"
#include "qboot/conformal_scale.hpp"
#include "qboot/context.hpp"
#include "qboot/hor_formula.hpp"
#include "qboot/hor_recursion.hpp"
#include "qboot/mp/rational.hpp"
#include "qboot/mp/real.hpp"
#include "qboot/my_filesystem.hpp"
#include "qboot/polynomial_program.hpp"
#include "qboot/primary_op.hpp"
#include "qboot/scale_factor.hpp"
#include "qboot/sdpb_input.hpp"
#include "qboot/task_queue.hpp"
#include "qboot/xml_input.hpp"
#endif  // QBOOT_B200_QBOOT_HPP_
