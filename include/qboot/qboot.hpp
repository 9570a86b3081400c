// qboot_b200 façade -- umbrella header: a program written for the reference says `#include "qboot/qboot.hpp"` and gets
// the whole call surface.  Listed by layer, bottom up; paths are relative to this directory.
#ifndef QBOOT_B200_QBOOT_HPP_
#define QBOOT_B200_QBOOT_HPP_
// numbers (records of the engine behind MPFR-compatible classes) and the host pool
#include "mp/real.hpp"
#include "mp/rational.hpp"
#include "task_queue.hpp"
#include "my_filesystem.hpp"
// containers and function spaces
#include "algebra/matrix.hpp"
#include "algebra/polynomial.hpp"
#include "algebra/real_function.hpp"
#include "algebra/complex_function.hpp"
// the model description: operators, blocks, scale factors
#include "primary_op.hpp"
#include "block.hpp"
#include "scale_factor.hpp"
#include "conformal_scale.hpp"
// the engine-backed Context and the table entry points
#include "context.hpp"
#include "hor_formula.hpp"
#include "hor_recursion.hpp"
// from crossing equations to SDPB input
#include "bootstrap_equation.hpp"
#include "polynomial_program.hpp"
#include "sdpb_input.hpp"
#include "xml_input.hpp"
#endif  // QBOOT_B200_QBOOT_HPP_
