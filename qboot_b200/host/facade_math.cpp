// qboot_b200 host library -- the numerical parts of the C++ façade that run on the host: dense linear algebra on
// mp::real, polynomial helpers and the scale factor of a continuous sector.
//
// Parity: every routine performs the reference's operations in the reference's order, each rounded once at
// mp::global_prec, so its outputs are the reference's bits.  Cited per function.
#include <sstream>
#include <utility>

#include "qboot/algebra/matrix.hpp"
#include "qboot/algebra/polynomial.hpp"
#include "qboot/conformal_scale.hpp"
#include "qboot/context.hpp"
#include "qboot/scale_factor.hpp"

namespace qboot::mp
{
	mpfr_prec_t global_prec = 1000;      // src/mp/real.cpp:5
	mpfr_rnd_t global_rnd = MPFR_RNDN;   // src/mp/real.cpp:6 (the only mode implemented)
}  // namespace qboot::mp

namespace qboot::algebra
{
	using mp::real;
	namespace
	{
		// row t += x * row f, from column c0 on
		void add_row(Matrix<real>& m, uint32_t f, uint32_t t, const real& x, uint32_t c0 = 0)
		{
			for (uint32_t c = c0; c < m.column(); ++c) m.at(t, c) += x * m.at(f, c);
		}
		void swap_row(Matrix<real>& m, uint32_t r1, uint32_t r2, uint32_t c0 = 0)
		{
			for (uint32_t c = c0; c < m.column(); ++c) m.at(r1, c).swap(m.at(r2, c));
		}
		void scale_row(Matrix<real>& m, uint32_t r, const real& x, uint32_t c0 = 0)
		{
			for (uint32_t c = c0; c < m.column(); ++c) m.at(r, c) *= x;
		}
		uint32_t pivot_row(const Matrix<real>& m, uint32_t j)
		{
			uint32_t p = j;
			for (uint32_t i = j + 1; i < m.row(); ++i)
				if (mp::cmpabs(m.at(p, j), m.at(i, j)) < 0) p = i;
			return p;
		}
	}  // namespace
	// Gaussian elimination with partial pivoting, the pivot row normalised first (src/algebra/matrix.cpp:38-64)
	real determinant(Matrix<real>&& m)
	{
		assert(m.is_square());
		const uint32_t n = m.row();
		if (n == 0) return real(1);
		if (n == 1) return m.at(0, 0);
		if (n == 2) return m.at(0, 0) * m.at(1, 1) - m.at(0, 1) * m.at(1, 0);
		real det(1);
		for (uint32_t j = 0; j < n; ++j)
		{
			const uint32_t p = pivot_row(m, j);
			if (p != j)
			{
				det.negate();
				swap_row(m, p, j, j);
			}
			det *= m.at(j, j);
			if (det.iszero()) break;
			scale_row(m, j, 1 / m.at(j, j), j);
			for (uint32_t i = j + 1; i < n; ++i) add_row(m, j, i, -m.at(i, j), j);
		}
		return det;
	}
	// Gauss-Jordan: forward elimination on [m | 1], then back substitution on the right half (src/algebra/matrix.cpp:65-86)
	Matrix<real> inverse(Matrix<real>&& m)
	{
		assert(m.is_square());
		const uint32_t n = m.row();
		Matrix<real> inv = Matrix<real>::constant(real(1), n);
		for (uint32_t j = 0; j < n; ++j)
		{
			const uint32_t p = pivot_row(m, j);
			swap_row(inv, p, j);
			swap_row(m, p, j, j);
			const real t = 1 / m.at(j, j);
			scale_row(inv, j, t);
			scale_row(m, j, t, j);
			for (uint32_t i = j + 1; i < n; ++i)
			{
				const real f = -m.at(i, j);
				add_row(inv, j, i, f);
				add_row(m, j, i, f, j);
			}
		}
		for (uint32_t j = n; j-- > 0;)
			for (uint32_t r = j; r-- > 0;) add_row(inv, j, r, -m.at(r, j));
		return inv;
	}
	// L(i, j) = (a(i, j) - sum_{k < j} L(i, k) L(j, k)) / L(j, j), sqrt on the diagonal (src/algebra/matrix.cpp:87-104)
	Matrix<real> cholesky_decomposition(const Matrix<real>& a)
	{
		assert(a.is_square());
		const uint32_t n = a.row();
		Matrix<real> L(n, n);
		for (uint32_t i = 0; i < n; ++i)
			for (uint32_t j = 0; j <= i; ++j)
			{
				real s{};
				for (uint32_t k = 0; k < j; ++k) s += L.at(i, k) * L.at(j, k);
				s = a.at(i, j) - s;
				L.at(i, j) = i == j ? mp::sqrt(s) : s / L.at(j, j);
			}
		return L;
	}
	// row by row: X(i, i) = 1 / L(i, i), X(i, j) = -(sum_{j <= k < i} L(i, k) X(k, j)) / L(i, i) (src/algebra/matrix.cpp:105-123)
	Matrix<real> lower_triangular_inverse(const Matrix<real>& l)
	{
		assert(l.is_square());
		const uint32_t n = l.row();
		Matrix<real> x(n, n);
		for (uint32_t i = 0; i < n; ++i)
		{
			x.at(i, i) = 1 / l.at(i, i);
			for (uint32_t j = 0; j < i; ++j)
			{
				real s{};
				for (uint32_t k = j; k < i; ++k) s += l.at(i, k) * x.at(k, j);
				x.at(i, j) = -s / l.at(i, i);
			}
		}
		return x;
	}
	// inverse of the Vandermonde matrix V(i, j) = points[i]^j built by repeated multiplication (src/algebra/polynomial.cpp:182-193)
	Matrix<real> interpolation_matrix(const Vector<real>& points)
	{
		assert(points.size() > 0);
		const uint32_t n = points.size();
		Matrix<real> v(n, n);
		for (uint32_t i = 0; i < n; ++i)
		{
			v.at(i, 0) = 1;
			for (uint32_t j = 1; j < n; ++j) v.at(i, j) = points[i] * v.at(i, j - 1);
		}
		return inverse(std::move(v));
	}
	Polynomial polynomial_interpolate(const Vector<real>& vals, const Matrix<real>& interpolation_matrix)
	{
		assert(vals.size() == interpolation_matrix.row() && vals.size() > 0 && interpolation_matrix.is_square());
		auto coeffs = dot(interpolation_matrix, vals);
		return Polynomial(std::move(coeffs));
	}
	std::string Polynomial::str(char var) const
	{
		if (iszero()) return "0";
		std::ostringstream out;
		bool first = true;
		for (int32_t i = 0; i <= degree(); ++i)
		{
			const real& c = at(uint32_t(i));
			if (c.iszero()) continue;
			if (!first) out << (mp::sgn(c) < 0 ? " - " : " + ");
			const real a = first ? c : mp::abs(c);
			if (i == 0)
				out << a;
			else
			{
				if (a != 1) out << a << " * ";
				out << var;
				if (i > 1) out << "^" << i;
			}
			first = false;
		}
		return out.str();
	}
}  // namespace qboot::algebra

namespace qboot
{
	using algebra::Matrix, algebra::Polynomial, algebra::Vector;
	using mp::rational, mp::real;

	ScaleFactor::~ScaleFactor() = default;
	TrivialScale::~TrivialScale() = default;
	// the constant factor: one sample (the reference puts it at x = 1 in the list and at 0 in the single query), value 1
	uint32_t TrivialScale::max_degree() const { return 0; }
	mp::real TrivialScale::eval(const mp::real&) const { return mp::real(1); }
	mp::real TrivialScale::sample_point(uint32_t) const { return mp::real(0); }
	ScaleFactor::Reals TrivialScale::sample_points() const { return Reals{mp::real(1)}; }
	ScaleFactor::Reals TrivialScale::sample_scalings() const { return Reals{mp::real(1)}; }
	ScaleFactor::Polys TrivialScale::bilinear_bases() const { return Polys{algebra::Polynomial(mp::real(1))}; }
	ConformalScale::~ConformalScale() = default;

	namespace
	{
		// poles of a block in Delta, each family listed from the largest down (src/conformal_scale.cpp:84-130):
		//   family 1: -(spin + k - 1)   family 2: epsilon - (k - 1)   family 3: 2 epsilon + 1 + spin - k  (k <= spin)
		// with k = 1, 2, 3, ... -- families 1 and 3 keep only even k unless odd poles are included.
		struct PoleFamily
		{
			int type;
			uint32_t spin, k;
			bool odd;
			rational eps;
			PoleFamily(int t, uint32_t s, const rational& e, bool include_odd) : type(t), spin(s), k(t != 2 && !include_odd ? 2 : 1), odd(include_odd), eps(e) {}
			bool alive() const { return type != 3 || k <= spin; }
			rational value() const
			{
				if (type == 1) return rational(-int64_t(spin + k - 1));
				if (type == 2) return eps - rational(int64_t(k - 1));
				return 2 * eps + rational(int64_t(1 + spin - k));
			}
			void advance() { k += (type == 2 || odd) ? 1 : 2; }
		};
		// largest-first merge of the three families; ties take the earlier family (1 before 2 before 3)
		rational next_pole(PoleFamily (&f)[3])
		{
			// the reference merges (1 with 2) with 3, taking the left stream when its head is >= the right one's
			auto pick12 = [&]() -> int { return !f[1].alive() || (f[0].alive() && f[0].value() >= f[1].value()) ? 0 : 1; };
			const int l = pick12();
			const bool left_alive = f[0].alive() || f[1].alive();
			const int pick = !f[2].alive() || (left_alive && f[l].value() >= f[2].value()) ? l : 2;
			rational v = f[pick].value();
			f[pick].advance();
			return v;
		}
		// A(i, j) = band[i + j];  rows of (chol A)^-1 are the orthonormal polynomials   (src/conformal_scale.cpp:11-18)
		Matrix<real> hankel_orthonormalise(const Vector<real>& band)
		{
			const uint32_t n = (band.size() + 1) / 2;
			Matrix<real> a(n, n);
			for (uint32_t i = 0; i < n; ++i)
				for (uint32_t j = 0; j < n; ++j) a.at(i, j) = band[i + j];
			return algebra::lower_triangular_inverse(algebra::cholesky_decomposition(a));
		}
		// int_0^inf base^x x^n dx = n! / (-ln base)^(n+1), n = 0..order   (src/conformal_scale.cpp:20-31)
		// (ln base is the same correctly rounded number wherever it is evaluated: the caller computes it once per scale factor)
		Vector<real> moments_no_pole(uint32_t order, const real& lnb)
		{
			Vector<real> m(order + 1);
			const real inv = -1 / lnb;
			m[0] = inv;
			for (uint32_t j = 1; j <= order; ++j) m[j] = m[j - 1] * j * inv;
			return m;
		}
		// int_0^inf base^x x^n / (x - pole) dx for n = 0..order, pole <= 0: with E = base^pole Gamma(0, pole ln base),
		//   m[0] = E,   m[j] = sum_{i < j} (j - i - 1)! pole^i (-1/ln base)^(j - i) + E pole^j      (src/conformal_scale.cpp:33-70)
		Vector<real> moments_simple_pole(uint32_t order, const real& base, const real& lnb, const real& pole)
		{
			const real e = pole == 0 ? real(mp::global_prec) : mp::pow(base, pole) * mp::gamma_inc(real(0), pole * lnb);
			Vector<real> m(order + 1);
			m[0] = e;
			real run{}, tail = pole * e, fact(1);
			const real inv = -1 / lnb;
			real inv_pow = inv;
			for (uint32_t j = 1; j <= order; ++j)
			{
				run = fact * inv_pow + run * pole;
				m[j] = run + tail;
				if (j < order)
				{
					tail *= pole;
					inv_pow *= inv;
					fact *= j;
				}
			}
			return m;
		}
		// w_i = 1 / prod_{j != i} (p_i - p_j)   (src/conformal_scale.cpp:72-82)
		Vector<real> partial_fraction_weights(const Vector<real>& poles)
		{
			Vector<real> w(poles.size());
			for (uint32_t i = 0; i < poles.size(); ++i)
			{
				w[i] = 1;
				for (uint32_t j = 0; j < poles.size(); ++j)
					if (i != j) w[i] *= poles[i] - poles[j];
				w[i] = 1 / w[i];
			}
			return w;
		}
	}  // namespace

	// src/conformal_scale.cpp:136-187
	void ConformalScale::_set_bilinear_bases() &
	{
		const uint32_t deg = bilinear_bases_.size() - 1;
		if (end_.has_value())
		{
			// finite range: monomials
			for (uint32_t i = 0; i <= deg; ++i) bilinear_bases_.at(i) = Polynomial(i);
			return;
		}
		const uint32_t np = poles_.size();
		Vector<real> shifted(np);
		for (uint32_t i = 0; i < np; ++i) shifted[i] = poles_[i] - gap_;
		{
			// a pole at x = 0 and double poles are moved by 1e-10 (2e-10 for the second pole at zero); this only touches the bases
			const real tiny("1e-10");
			uint32_t i = 0;
			if (i < np && shifted[i] == 0)
			{
				shifted[i++] -= tiny;
				if (i < np && shifted[i] == 0) shifted[i++] -= 2 * tiny;
			}
			for (; i < np; ++i)
				if (i > 0 && shifted[i] == shifted[i - 1]) shifted[i] -= tiny;
		}
		const Vector<real> weight = partial_fraction_weights(shifted);
		const real base = 4 * rho_, lnb = mp::log(base);
		Vector<real> moments(2 * deg + 1);
		if (np == 0)
			moments = moments_no_pole(2 * deg, lnb);
		else
			for (uint32_t i = 0; i < np; ++i) moments += mul_scalar(weight[i], moments_simple_pole(2 * deg, base, lnb, shifted[i]));
		moments *= mp::pow(base, gap_);
		Matrix<real> q = hankel_orthonormalise(moments);
		for (uint32_t i = 0; i <= deg; ++i)
		{
			Vector<real> row(i + 1);
			for (uint32_t j = 0; j <= i; ++j) row[j] = q.at(i, j);
			bilinear_bases_.at(i) = Polynomial(std::move(row));
		}
	}
	ConformalScale::ConformalScale(uint32_t cutoff, uint32_t spin, uint32_t lambda, const rational& epsilon, const real& rho,
	                               bool include_odd, const real& gap, const std::optional<real>& end)
	    : odd_included_(include_odd), spin_(spin), lambda_(lambda), epsilon_(epsilon), rho_(rho), gap_(gap), end_(end), poles_(cutoff),
	      bilinear_bases_((cutoff + lambda) / 2 + 1)
	{
		PoleFamily fam[3] = {PoleFamily(1, spin, epsilon, include_odd), PoleFamily(2, spin, epsilon, include_odd), PoleFamily(3, spin, epsilon, include_odd)};
		for (uint32_t i = 0; i < cutoff; ++i) poles_[i] = real(next_pole(fam));
		_set_bilinear_bases();
	}
	ConformalScale::ConformalScale(uint32_t cutoff, uint32_t spin, const Context& c, bool include_odd, const real& gap,
	                               const std::optional<real>& end)
	    : ConformalScale(cutoff, spin, c.lambda(), c.epsilon(), c.rho(), include_odd, gap, end)
	{
	}
	ConformalScale::ConformalScale(const GeneralPrimaryOperator& op, const Context& c, bool include_odd)
	    : ConformalScale(op.num_poles(), op.spin(), c, include_odd, op.lower_bound(), op.upper_bound_safe())
	{
	}
	ConformalScale::ConformalScale(_DeferBases, const GeneralPrimaryOperator& op, const Context& c, bool include_odd)
	    : odd_included_(include_odd), spin_(op.spin()), lambda_(c.lambda()), epsilon_(c.epsilon()), rho_(c.rho()), gap_(op.lower_bound()),
	      end_(op.upper_bound_safe()), poles_(op.num_poles()), bilinear_bases_((op.num_poles() + c.lambda()) / 2 + 1), bases_ready_(false)
	{
		PoleFamily fam[3] = {PoleFamily(1, spin_, epsilon_, include_odd), PoleFamily(2, spin_, epsilon_, include_odd), PoleFamily(3, spin_, epsilon_, include_odd)};
		for (uint32_t i = 0; i < poles_.size(); ++i) poles_[i] = real(next_pole(fam));
	}
	// include/qboot/conformal_scale.hpp:66-72
	real ConformalScale::eval_d(const real& delta) const
	{
		real chi = mp::pow(4 * rho_, delta);
		if (end_.has_value()) chi *= mp::pow(get_x(delta) + end_.value(), -int32_t(max_degree()));
		for (const auto& p : poles_) chi /= delta - p;
		return chi;
	}
	// include/qboot/conformal_scale.hpp:74-85
	real ConformalScale::get_delta(const real& x) const
	{
		if (!end_.has_value()) return x + gap_;
		return end_.value() * (x + gap_) / (x + end_.value());
	}
	real ConformalScale::get_x(const real& delta) const
	{
		if (!end_.has_value()) return delta - gap_;
		return end_.value() * (delta - gap_) / (end_.value() - delta);
	}
	Vector<real> ConformalScale::sample_scalings() const
	{
		Vector<real> s = sample_points();
		for (auto& x : s) x = eval(x);
		return s;
	}
	// include/qboot/conformal_scale.hpp:105-116 (the single-point form divides each point on its own; both round alike)
	real ConformalScale::sample_point(uint32_t degree, uint32_t k)
	{
		return mp::pow(mp::const_pi() * (-1 + 4 * int32_t(k)), 2uL) / ((-64 * int32_t(degree + 1)) * mp::log(3 - mp::sqrt(8)));
	}
	Vector<real> ConformalScale::sample_points(uint32_t degree)
	{
		Vector<real> x(degree + 1);
		const real pi = mp::const_pi();
		for (uint32_t k = 0; k <= degree; ++k) x[k] = mp::pow(pi * (-1 + 4 * int32_t(k)), 2uL);
		x /= (-64 * int32_t(degree + 1)) * mp::log(3 - mp::sqrt(8));
		return x;
	}
}  // namespace qboot
