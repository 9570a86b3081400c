// qboot_b200 façade -- Vector / Matrix containers with the reference's call surface
// (include/qboot/algebra/matrix.hpp, src/algebra/matrix.cpp).
//
// Host-side glue only: they carry model data, scale factors and the pieces of a polynomial matrix program between the
// planner, the C ABI and the writers.  The numerical kernels of the linear algebra a ConformalScale needs (Cholesky
// factor and triangular inverse of the Hankel moment matrix, src/algebra/matrix.cpp:87-123; inverse and determinant
// with partial pivoting, :38-86) perform the reference's operations in the reference's order, so results agree bit for bit.
#ifndef QBOOT_B200_ALGEBRA_MATRIX_HPP_
#define QBOOT_B200_ALGEBRA_MATRIX_HPP_

#include <cassert>
#include <cstdint>
#include <initializer_list>
#include <ostream>
#include <type_traits>
#include <utility>
#include <vector>

#include "qboot/mp/real.hpp"

namespace qboot::algebra
{
	// move-only, explicit clone(): the ownership convention of the reference's containers
	template <class Ring>
	class Vector
	{
		std::vector<Ring> v_;

	public:
		void _reset() && { std::vector<Ring>{}.swap(v_); }
		Vector() = default;
		explicit Vector(uint32_t len) : v_(len) {}
		Vector(uint32_t len, const Ring& val)
		{
			v_.reserve(len);
			for (uint32_t i = 0; i < len; ++i) v_.push_back(val.clone());
		}
		Vector(std::initializer_list<Ring> items)
		{
			v_.reserve(items.size());
			for (const auto& t : items) v_.push_back(t.clone());
		}
		Vector(Vector&&) noexcept = default;
		Vector& operator=(Vector&&) & noexcept = default;
		Vector(const Vector&) = delete;
		Vector& operator=(const Vector&) = delete;
		~Vector() = default;
		[[nodiscard]] Ring& at(uint32_t i) & { return v_[i]; }
		[[nodiscard]] const Ring& at(uint32_t i) const& { return v_[i]; }
		[[nodiscard]] Ring& operator[](uint32_t i) & { return v_[i]; }
		[[nodiscard]] const Ring& operator[](uint32_t i) const& { return v_[i]; }
		[[nodiscard]] uint32_t size() const noexcept { return uint32_t(v_.size()); }
		[[nodiscard]] const Ring* begin() const& noexcept { return v_.data(); }
		[[nodiscard]] const Ring* end() const& noexcept { return v_.data() + v_.size(); }
		[[nodiscard]] Ring* begin() & noexcept { return v_.data(); }
		[[nodiscard]] Ring* end() & noexcept { return v_.data() + v_.size(); }
		[[nodiscard]] Vector clone() const
		{
			Vector r;
			r.v_.reserve(v_.size());
			for (const auto& x : v_) r.v_.push_back(x.clone());
			return r;
		}
		void swap(Vector& o) & { v_.swap(o.v_); }
		void negate() &
		{
			for (auto& x : v_) x.negate();
		}
		[[nodiscard]] bool iszero() const noexcept
		{
			for (const auto& x : v_)
				if (!x.iszero()) return false;
			return true;
		}
		[[nodiscard]] auto norm() const
		{
			auto s = v_.at(0).norm();
			for (size_t i = 1; i < v_.size(); ++i) s += v_[i].norm();
			return s;
		}
		[[nodiscard]] auto abs() const { return mp::sqrt(norm()); }
		Vector operator+() const& { return clone(); }
		Vector operator+() && { return std::move(*this); }
		Vector operator-() const&
		{
			Vector r = clone();
			r.negate();
			return r;
		}
		Vector operator-() &&
		{
			negate();
			return std::move(*this);
		}
		Vector& operator+=(const Vector& o) &
		{
			assert(size() == o.size());
			for (size_t i = 0; i < v_.size(); ++i) v_[i] += o.v_[i];
			return *this;
		}
		Vector& operator-=(const Vector& o) &
		{
			assert(size() == o.size());
			for (size_t i = 0; i < v_.size(); ++i) v_[i] -= o.v_[i];
			return *this;
		}
		template <class R>
		Vector& operator*=(const R& c) &
		{
			for (auto& x : v_) x *= c;
			return *this;
		}
		template <class R>
		Vector& operator/=(const R& c) &
		{
			for (auto& x : v_) x /= c;
			return *this;
		}
		friend Vector operator+(const Vector& a, const Vector& b)
		{
			Vector r = a.clone();
			r += b;
			return r;
		}
		friend Vector operator-(const Vector& a, const Vector& b)
		{
			Vector r = a.clone();
			r -= b;
			return r;
		}
		// c * v, elementwise mul_scalar(c, v[i]) = c * v[i]
		template <class R>
		friend Vector mul_scalar(const R& c, const Vector& v)
		{
			Vector r(v.size());
			for (uint32_t i = 0; i < v.size(); ++i) r.v_[i] = mul_scalar(c, v.v_[i]);
			return r;
		}
		template <class R>
		friend Vector mul_scalar(const R& c, Vector&& v)
		{
			return mul_scalar(c, static_cast<const Vector&>(v));
		}
		template <class R>
		friend Vector operator/(const Vector& v, const R& c)
		{
			Vector r = v.clone();
			r /= c;
			return r;
		}
		friend bool operator==(const Vector& a, const Vector& b)
		{
			if (a.size() != b.size()) return false;
			for (uint32_t i = 0; i < a.size(); ++i)
				if (!(a.v_[i] == b.v_[i])) return false;
			return true;
		}
		friend bool operator!=(const Vector& a, const Vector& b) { return !(a == b); }
		template <class Char, class Traits>
		friend std::basic_ostream<Char, Traits>& operator<<(std::basic_ostream<Char, Traits>& out, const Vector& v)
		{
			out << "[";
			for (uint32_t i = 0; i < v.size(); ++i) out << (i ? ", " : "") << v.v_[i];
			return out << "]";
		}
	};
	template <class Ring>
	class Matrix
	{
		Vector<Ring> a_;
		uint32_t row_ = 0, col_ = 0;

	public:
		void _reset() &&
		{
			std::move(a_)._reset();
			row_ = col_ = 0;
		}
		Matrix() = default;
		Matrix(uint32_t r, uint32_t c) : a_(r * c), row_(r), col_(c) {}
		Matrix(Matrix&&) noexcept = default;
		Matrix& operator=(Matrix&&) & noexcept = default;
		Matrix(const Matrix&) = delete;
		Matrix& operator=(const Matrix&) = delete;
		~Matrix() = default;
		[[nodiscard]] Ring& at(uint32_t r, uint32_t c) & { return a_[r * col_ + c]; }
		[[nodiscard]] const Ring& at(uint32_t r, uint32_t c) const& { return a_[r * col_ + c]; }
		[[nodiscard]] uint32_t row() const noexcept { return row_; }
		[[nodiscard]] uint32_t column() const noexcept { return col_; }
		[[nodiscard]] bool is_square() const noexcept { return row_ == col_; }
		[[nodiscard]] const Vector<Ring>& _flat() const { return a_; }
		[[nodiscard]] Matrix clone() const
		{
			Matrix m;
			m.a_ = a_.clone();
			m.row_ = row_;
			m.col_ = col_;
			return m;
		}
		void swap(Matrix& o) &
		{
			a_.swap(o.a_);
			std::swap(row_, o.row_);
			std::swap(col_, o.col_);
		}
		void negate() & { a_.negate(); }
		[[nodiscard]] bool iszero() const noexcept { return a_.iszero(); }
		[[nodiscard]] auto norm() const { return a_.norm(); }
		[[nodiscard]] auto abs() const { return mp::sqrt(norm()); }
		void transpose() &
		{
			Matrix t(col_, row_);
			for (uint32_t r = 0; r < row_; ++r)
				for (uint32_t c = 0; c < col_; ++c) t.at(c, r).swap(at(r, c));
			swap(t);
		}
		// n x n with c on the diagonal
		static Matrix constant(const Ring& c, uint32_t n)
		{
			Matrix m(n, n);
			for (uint32_t i = 0; i < n; ++i) m.at(i, i) = c.clone();
			return m;
		}
		Matrix operator+() const& { return clone(); }
		Matrix operator+() && { return std::move(*this); }
		Matrix operator-() const&
		{
			Matrix m = clone();
			m.negate();
			return m;
		}
		Matrix operator-() &&
		{
			negate();
			return std::move(*this);
		}
		// v^t M v, accumulated row-major with fma as the reference does (include/qboot/algebra/matrix.hpp:351-359)
		[[nodiscard]] Ring inner_product(const Vector<Ring>& v) const
		{
			assert(is_square() && row_ == v.size());
			mp::real s{};
			for (uint32_t r = 0; r < row_; ++r)
				for (uint32_t c = 0; c < row_; ++c) mp::fma(s, mp::mul(v[r], v[c]), at(r, c), s);
			return s;
		}
		Matrix& operator+=(const Matrix& o) &
		{
			a_ += o.a_;
			return *this;
		}
		Matrix& operator-=(const Matrix& o) &
		{
			a_ -= o.a_;
			return *this;
		}
		template <class R>
		Matrix& operator*=(const R& c) &
		{
			a_ *= c;
			return *this;
		}
		template <class R>
		Matrix& operator/=(const R& c) &
		{
			a_ /= c;
			return *this;
		}
		friend Matrix operator+(const Matrix& a, const Matrix& b)
		{
			Matrix m = a.clone();
			m += b;
			return m;
		}
		friend Matrix operator-(const Matrix& a, const Matrix& b)
		{
			Matrix m = a.clone();
			m -= b;
			return m;
		}
		template <class R>
		friend Matrix mul_scalar(const R& c, const Matrix& m)
		{
			Matrix r;
			r.a_ = mul_scalar(c, m.a_);
			r.row_ = m.row_;
			r.col_ = m.col_;
			return r;
		}
		template <class R>
		friend Matrix mul_scalar(const R& c, Matrix&& m)
		{
			return mul_scalar(c, static_cast<const Matrix&>(m));
		}
		template <class R>
		friend Matrix operator/(const Matrix& m, const R& c)
		{
			Matrix r = m.clone();
			r /= c;
			return r;
		}
		template <class R>
		friend Matrix operator/(Matrix&& m, const R& c)
		{
			m /= c;
			return std::move(m);
		}
		friend bool operator==(const Matrix& a, const Matrix& b) { return a.row_ == b.row_ && a.col_ == b.col_ && a.a_ == b.a_; }
		friend bool operator!=(const Matrix& a, const Matrix& b) { return !(a == b); }
		template <class Char, class Traits>
		friend std::basic_ostream<Char, Traits>& operator<<(std::basic_ostream<Char, Traits>& out, const Matrix& m)
		{
			out << "[";
			for (uint32_t r = 0; r < m.row_; ++r)
			{
				out << (r ? ", [" : "[");
				for (uint32_t c = 0; c < m.col_; ++c) out << (c ? ", " : "") << m.at(r, c);
				out << "]";
			}
			return out << "]";
		}
	};

	// z[i] = sum_j x[j] * M[j][i], the first product assigned and the rest added in order (include/qboot/algebra/matrix.hpp:444-456)
	template <class R1, class R2>
	auto dot(const Vector<R1>& x, const Matrix<R2>& m)
	{
		assert(x.size() == m.row());
		Vector<decltype(mul(x[0], m.at(0, 0)))> z(m.column());
		for (uint32_t i = 0; i < m.column(); ++i)
		{
			z[i] = mul(x[0], m.at(0, i));
			for (uint32_t j = 1; j < m.row(); ++j) z[i] += mul(x[j], m.at(j, i));
		}
		return z;
	}
	template <class R1, class R2>
	auto dot(const Matrix<R1>& m, const Vector<R2>& x)
	{
		assert(x.size() == m.column());
		Vector<decltype(mul_scalar(m.at(0, 0), x[0]))> z(m.row());
		for (uint32_t i = 0; i < m.row(); ++i)
		{
			z[i] = mul_scalar(m.at(i, 0), x[0]);
			for (uint32_t j = 1; j < m.column(); ++j) z[i] += mul_scalar(m.at(i, j), x[j]);
		}
		return z;
	}
	template <class R>
	Matrix<R> dot(const Matrix<R>& a, const Matrix<R>& b)
	{
		assert(a.column() == b.row());
		Matrix<R> z(a.row(), b.column());
		for (uint32_t i = 0; i < a.row(); ++i)
			for (uint32_t j = 0; j < b.column(); ++j)
			{
				z.at(i, j) = mul(a.at(i, 0), b.at(0, j));
				for (uint32_t k = 1; k < a.column(); ++k) z.at(i, j) += mul(a.at(i, k), b.at(k, j));
			}
		return z;
	}

	// implemented in qboot_b200/host/facade.cpp, operation order of src/algebra/matrix.cpp:38-123
	mp::real determinant(Matrix<mp::real>&& mat);
	inline mp::real determinant(const Matrix<mp::real>& mat) { return determinant(mat.clone()); }
	Matrix<mp::real> inverse(Matrix<mp::real>&& mat);
	inline Matrix<mp::real> inverse(const Matrix<mp::real>& mat) { return inverse(mat.clone()); }
	// lower triangular L with L L^t = mat, for a positive definite mat
	Matrix<mp::real> cholesky_decomposition(const Matrix<mp::real>& mat);
	Matrix<mp::real> lower_triangular_inverse(const Matrix<mp::real>& mat);
}  // namespace qboot::algebra

#endif  // QBOOT_B200_ALGEBRA_MATRIX_HPP_
