// qboot_b200 host library -- the façade's link to the CUDA engine: an RAII owner of one qb_handle and the device-side
// state of a packed program.  Everything below this file is the C ABI of include/qboot_b200.h; nothing here knows CUDA.
#pragma once

#include <cstdint>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "qboot/mp/rational.hpp"
#include "qboot_b200.h"

namespace qboot::b200
{
	// one engine handle with a Context's constants uploaded; calls are serialised by `mu`
	class Engine
	{
	public:
		qb_handle* h = nullptr;
		int prec = 0, words = 0, device = 0;
		uint32_t n_max = 0, lambda = 0;
		int dim_mixed = 0, dim_sym[2] = {0, 0};
		std::recursive_mutex mu;
		uint64_t epoch = 0;       // generation of the constraint-block arena (bumped by every assemble)
		uint64_t pack_epoch = 0;  // generation of the packed arrays (bumped by every pack)
		Engine(int prec_bits, uint32_t n_max_, uint32_t lambda_, const mp::rational& epsilon);
		Engine(const Engine&) = delete;
		Engine& operator=(const Engine&) = delete;
		~Engine();
		// throw std::runtime_error with the engine's message unless rc == 0 (or rc >= 0 when nonneg_ok)
		int check(int rc, const char* what, bool nonneg_ok = false) const;
	};
	// the packed d_B / d_c arrays of one PolynomialProgram::create_input, and (once asked for) their text
	struct PackedProgram
	{
		std::shared_ptr<Engine> engine;
		uint64_t pack_epoch = 0;
		int n_free = 0, n_blocks = 0;
		// text of all constraints as produced by qb_pmp_text (pinned host memory), and where each piece starts
		char* text = nullptr;
		int64_t text_bytes = 0;
		size_t text_cap = 0;  // bytes of the pinned buffer behind `text`
		std::vector<int64_t> sizes, offsets;  // 2 per block: d_B text, d_c text
		int text_digits = 0;
		std::mutex mu;
		void require_current() const
		{
			if (engine->pack_epoch != pack_epoch)
				throw std::logic_error("qboot_b200: this SDPBInput's device arrays were replaced by a later create_input on the same Context");
		}
		void ensure_text(int ndigits);
		~PackedProgram();
	};
	// pinned host buffers for bulk transfers (qb_pinned_alloc)
	struct Pinned
	{
		void* p = nullptr;
		explicit Pinned(size_t bytes);
		~Pinned();
		Pinned(const Pinned&) = delete;
		Pinned& operator=(const Pinned&) = delete;
	};
	// Page-locking hundreds of megabytes costs more than filling them (~0.1 s for the 400 MB of text of a mixed-correlator
	// program), so the process keeps the largest released buffer for the next taker.
	char* pinned_acquire(size_t bytes, size_t& capacity);
	void pinned_release(char* p, size_t bytes);
	// n numbers (records, host) -> their SDPB text lines, formatted on the device (qb_dec_format) with the host formatter
	// for the few the device declines; appended to `out`
	// (ends, when given, receives the offset in `out` just past each number's line)
	void format_numbers(Engine& e, int ndigits, const uint32_t* recs, size_t n, std::string& out, std::vector<size_t>* ends = nullptr);
}  // namespace qboot::b200
