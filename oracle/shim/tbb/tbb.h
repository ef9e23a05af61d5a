// oracle/shim/tbb/tbb.h  --  TEST INFRASTRUCTURE ONLY (not part of the product).
//
// Minimal header-only stand-in for the subset of oneTBB that the reference's
// permanent / sampling sources use, so that those sources can be compiled
// UNCHANGED from /root/reference into oracle/_ref/ (TBB itself is not installed
// in this image).  It is an original implementation on top of <thread>; it is
// not derived from oneTBB.
//
// Provided: scalable_{malloc,free,aligned_malloc,aligned_free},
//   tbb::blocked_range, tbb::parallel_for (range and index forms),
//   tbb::parallel_invoke, tbb::task_group, tbb::combinable, tbb::spin_mutex,
//   tbb::cache_aligned_allocator, tbb::scalable_allocator,
//   tbb::concurrent_vector, tbb::tick_count.
//
// Threading model: one persistent pool of (PQB_ORACLE_THREADS or
// hardware_concurrency) workers; a top-level parallel_for hands out indices
// through an atomic counter; nested parallel constructs run inline on the
// calling worker (this is what a saturated TBB arena effectively does).
#pragma once

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

// ---------------------------------------------------------------- allocator
inline void* scalable_malloc(size_t n) { return std::malloc(n ? n : 1); }
inline void scalable_free(void* p) { std::free(p); }
inline void* scalable_aligned_malloc(size_t n, size_t align) {
    void* p = nullptr;
    if (align < sizeof(void*)) align = sizeof(void*);
    if (posix_memalign(&p, align, n ? n : align) != 0) return nullptr;
    return p;
}
inline void scalable_aligned_free(void* p) { std::free(p); }

namespace tbb {

// ---------------------------------------------------------------- pool
namespace shim_detail {

inline int& nesting_depth() {
    static thread_local int depth = 0;
    return depth;
}

class Pool {
public:
    static Pool& instance() {
        // leaked on purpose: workers are detached and wait on cv_ forever; destroying a
        // condition_variable that still has waiters would block process exit
        static Pool* pool = new Pool();
        return *pool;
    }

    int size() const { return nthreads_; }

    // run body(i) for i in [0, count) on the pool; the caller participates
    void run(int64_t count, const std::function<void(int64_t)>& body) {
        if (count <= 0) return;
        if (nthreads_ <= 1 || count == 1 || nesting_depth() > 0) {
            nesting_depth()++;
            for (int64_t i = 0; i < count; i++) body(i);
            nesting_depth()--;
            return;
        }
        std::unique_lock<std::mutex> serial(serial_mutex_);  // one top-level job at a time
        {
            std::lock_guard<std::mutex> lk(m_);
            body_ = &body;
            count_ = count;
            next_.store(0);
            pending_ = nthreads_ - 1;
            generation_++;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> lk(m_);
        done_cv_.wait(lk, [&] { return pending_ == 0; });
        body_ = nullptr;
    }

private:
    Pool() {
        const char* env = std::getenv("PQB_ORACLE_THREADS");
        int n = env ? std::atoi(env) : (int)std::thread::hardware_concurrency();
        if (n < 1) n = 1;
        nthreads_ = n;
        for (int i = 1; i < nthreads_; i++) {
            workers_.emplace_back([this] { loop(); });
            workers_.back().detach();
        }
    }

    void work() {
        nesting_depth()++;
        for (;;) {
            int64_t i = next_.fetch_add(1);
            if (i >= count_) break;
            (*body_)(i);
        }
        nesting_depth()--;
    }

    void loop() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return generation_ != seen; });
                seen = generation_;
            }
            work();
            {
                std::lock_guard<std::mutex> lk(m_);
                pending_--;
            }
            done_cv_.notify_one();
        }
    }

    int nthreads_ = 1;
    std::vector<std::thread> workers_;
    std::mutex m_, serial_mutex_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int64_t)>* body_ = nullptr;
    int64_t count_ = 0;
    std::atomic<int64_t> next_{0};
    int pending_ = 0;
    uint64_t generation_ = 0;
};

}  // namespace shim_detail

// ---------------------------------------------------------------- ranges
template <typename T>
class blocked_range {
public:
    blocked_range(T b, T e, size_t grain = 1) : b_(b), e_(e), grain_(grain) {}
    T begin() const { return b_; }
    T end() const { return e_; }
    size_t size() const { return (size_t)(e_ - b_); }
    size_t grainsize() const { return grain_; }
    bool empty() const { return !(b_ < e_); }

private:
    T b_, e_;
    size_t grain_;
};

template <typename T, typename F>
void parallel_for(const blocked_range<T>& range, const F& f) {
    if (range.empty()) return;
    int64_t total = (int64_t)(range.end() - range.begin());
    int64_t pieces = (int64_t)shim_detail::Pool::instance().size() * 4;
    if (shim_detail::nesting_depth() > 0) pieces = 1;
    if (pieces > total) pieces = total;
    T begin = range.begin();
    shim_detail::Pool::instance().run(pieces, [&](int64_t p) {
        T lo = begin + (T)(total * p / pieces);
        T hi = begin + (T)(total * (p + 1) / pieces);
        if (lo < hi) f(blocked_range<T>(lo, hi));
    });
}

template <typename I, typename F>
void parallel_for(I first, I last, I step, const F& f) {
    if (!(first < last)) return;
    int64_t count = (int64_t)((last - first + step - 1) / step);
    shim_detail::Pool::instance().run(count, [&](int64_t k) { f((I)(first + (I)k * step)); });
}

template <typename I, typename F>
void parallel_for(I first, I last, const F& f) {
    parallel_for(first, last, (I)1, f);
}

template <typename... Fs>
void parallel_invoke(const Fs&... fs) {
    std::function<void()> tasks[] = {std::function<void()>(fs)...};
    int64_t count = (int64_t)sizeof...(Fs);
    shim_detail::Pool::instance().run(count, [&](int64_t k) { tasks[k](); });
}

// task_group: tasks run inline at run() time (deterministic, deadlock free)
class task_group {
public:
    template <typename F>
    void run(const F& f) { f(); }
    template <typename F>
    void run_and_wait(const F& f) { f(); }
    void wait() {}
};

// ---------------------------------------------------------------- combinable
template <typename T>
class combinable {
public:
    combinable() : init_([] { return T(); }) {}
    template <typename Init>
    explicit combinable(Init init) : init_(init) {}
    combinable(const combinable& other) : init_(other.init_) {
        std::lock_guard<std::mutex> lk(other.m_);
        for (auto& s : other.slots_) slots_.push_back(std::make_pair(s.first, std::make_unique<T>(*s.second)));
    }
    combinable& operator=(const combinable& other) {
        if (this == &other) return *this;
        std::lock_guard<std::mutex> lk(m_);
        std::lock_guard<std::mutex> lk2(other.m_);
        init_ = other.init_;
        slots_.clear();
        for (auto& s : other.slots_) slots_.push_back(std::make_pair(s.first, std::make_unique<T>(*s.second)));
        return *this;
    }

    T& local() {
        std::thread::id me = std::this_thread::get_id();
        std::lock_guard<std::mutex> lk(m_);
        for (auto& s : slots_)
            if (s.first == me) return *s.second;
        slots_.push_back(std::make_pair(me, std::make_unique<T>(init_())));
        return *slots_.back().second;
    }

    template <typename F>
    void combine_each(F f) {
        std::lock_guard<std::mutex> lk(m_);
        for (auto& s : slots_) f(*s.second);
    }

    template <typename F>
    T combine(F f) {
        std::lock_guard<std::mutex> lk(m_);
        if (slots_.empty()) return init_();
        T acc = *slots_[0].second;
        for (size_t i = 1; i < slots_.size(); i++) acc = f(acc, *slots_[i].second);
        return acc;
    }

    void clear() {
        std::lock_guard<std::mutex> lk(m_);
        slots_.clear();
    }

private:
    std::function<T()> init_;
    mutable std::mutex m_;
    std::vector<std::pair<std::thread::id, std::unique_ptr<T>>> slots_;
};

// ---------------------------------------------------------------- mutex
class spin_mutex {
public:
    spin_mutex() {}
    void lock() {
        while (flag_.test_and_set(std::memory_order_acquire)) {
        }
    }
    void unlock() { flag_.clear(std::memory_order_release); }
    class scoped_lock {
    public:
        scoped_lock() : m_(nullptr) {}
        explicit scoped_lock(spin_mutex& m) : m_(&m) { m_->lock(); }
        ~scoped_lock() {
            if (m_) m_->unlock();
        }
        void acquire(spin_mutex& m) {
            m_ = &m;
            m_->lock();
        }
        void release() {
            if (m_) m_->unlock();
            m_ = nullptr;
        }

    private:
        spin_mutex* m_;
    };

private:
    std::atomic_flag flag_ = ATOMIC_FLAG_INIT;
};

// ---------------------------------------------------------------- containers
template <typename T>
using cache_aligned_allocator = std::allocator<T>;
template <typename T>
using scalable_allocator = std::allocator<T>;

template <typename T, typename A = std::allocator<T>>
class concurrent_vector : public std::vector<T, A> {
public:
    using std::vector<T, A>::vector;
};

// ---------------------------------------------------------------- timing
class tick_count {
public:
    class interval_t {
    public:
        explicit interval_t(double s = 0.0) : s_(s) {}
        double seconds() const { return s_; }

    private:
        double s_;
    };
    static tick_count now() {
        tick_count t;
        t.t_ = std::chrono::steady_clock::now();
        return t;
    }
    friend interval_t operator-(const tick_count& a, const tick_count& b) {
        return interval_t(std::chrono::duration<double>(a.t_ - b.t_).count());
    }

private:
    std::chrono::steady_clock::time_point t_;
};

}  // namespace tbb
