// qboot_b200 façade -- host-side fan-out and the profiling hook (call surface of include/qboot/task_queue.hpp).
//
// The reference parallelises a PMP build with short-lived threads pulling task indices under a mutex
// (task_queue.hpp:33-74) and a condvar worker pool behind every Context (:113-163).  On this engine the table work
// runs on the GPU; what is left for host threads is the per-constraint scale-factor construction and file output, for
// which the same _parallel_evaluate entry point is kept (results in task order, independent of the thread count).
#ifndef QBOOT_B200_TASK_QUEUE_HPP_
#define QBOOT_B200_TASK_QUEUE_HPP_

#include <atomic>
#include <cstdint>
#include <exception>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <string_view>
#include <thread>
#include <type_traits>
#include <vector>

namespace qboot
{
	// run tasks[i]() on up to `p` threads; returns the results in order (void tasks: nothing)
	template <class T>
	std::vector<T> _parallel_evaluate(const std::vector<std::function<T()>>& tasks, uint32_t p)
	{
		std::vector<T> out(tasks.size());
		std::atomic<size_t> next{0};
		std::exception_ptr err;
		std::mutex mu;
		auto worker = [&] {
			for (size_t i; (i = next.fetch_add(1)) < tasks.size();)
			{
				try
				{
					out[i] = tasks[i]();
				}
				catch (...)
				{
					std::lock_guard<std::mutex> g(mu);
					if (!err) err = std::current_exception();
				}
			}
		};
		const size_t nth = std::max<size_t>(1, std::min<size_t>(p, tasks.size()));
		std::vector<std::thread> pool;
		for (size_t t = 1; t < nth; ++t) pool.emplace_back(worker);
		worker();
		for (auto& th : pool) th.join();
		if (err) std::rethrow_exception(err);
		return out;
	}
	inline void _parallel_evaluate(const std::vector<std::function<void()>>& tasks, uint32_t p)
	{
		std::vector<std::function<int()>> wrapped;
		wrapped.reserve(tasks.size());
		for (const auto& f : tasks)
			wrapped.emplace_back([&f] {
				f();
				return 0;
			});
		_parallel_evaluate(wrapped, p);
	}

	// profiling hook: on_begin / on_end around named pieces of work
	class _event_base
	{
	public:
		_event_base() = default;
		_event_base(const _event_base&) = default;
		_event_base& operator=(const _event_base&) = default;
		_event_base(_event_base&&) noexcept = default;
		_event_base& operator=(_event_base&&) noexcept = default;
		virtual ~_event_base();
		virtual void on_begin(std::string_view tag) = 0;
		virtual void on_end(std::string_view tag) = 0;
	};
	class _scoped_event
	{
		std::string tag_;
		_event_base* ev_;

	public:
		_scoped_event(std::string_view tag, const std::unique_ptr<_event_base>& ev) : tag_(tag), ev_(ev.get())
		{
			if (ev_) ev_->on_begin(tag_);
		}
		_scoped_event(const _scoped_event&) = delete;
		_scoped_event& operator=(const _scoped_event&) = delete;
		~_scoped_event()
		{
			if (ev_) ev_->on_end(tag_);
		}
	};
}  // namespace qboot

#endif  // QBOOT_B200_TASK_QUEUE_HPP_
