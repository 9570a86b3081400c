/* TEST INFRASTRUCTURE ONLY (oracle build) -- not part of the product.
 *
 * Minimal stand-in for <gmpxx.h>: the reference includes it but uses only the
 * iostream operators on raw mpz/mpq pointers (include/qboot/mp/integer.hpp:846-854,
 * include/qboot/mp/rational.hpp:449-459).  Implemented via the C get_str/set_str API.
 */
#ifndef QB_ORACLE_SHIM_GMPXX_H_
#define QB_ORACLE_SHIM_GMPXX_H_

#include <cstdlib>
#include <cstring>  // the real gmpxx.h pulls this in; the reference relies on it for std::strlen
#include <istream>
#include <ostream>
#include <string>

#include "gmp.h"

inline std::ostream& operator<<(std::ostream& s, mpz_srcptr z)
{
	char* p = mpz_get_str(nullptr, 10, z);
	s << p;
	std::free(p);
	return s;
}
inline std::ostream& operator<<(std::ostream& s, mpq_srcptr q)
{
	char* p = mpq_get_str(nullptr, 10, q);
	s << p;
	std::free(p);
	return s;
}
inline std::istream& operator>>(std::istream& s, mpz_ptr z)
{
	std::string t;
	s >> t;
	if (mpz_set_str(z, t.c_str(), 10) != 0) s.setstate(std::ios::failbit);
	return s;
}
inline std::istream& operator>>(std::istream& s, mpq_ptr q)
{
	std::string t;
	s >> t;
	if (mpq_set_str(q, t.c_str(), 10) != 0) s.setstate(std::ios::failbit);
	return s;
}

#endif /* QB_ORACLE_SHIM_GMPXX_H_ */
