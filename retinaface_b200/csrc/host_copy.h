// host_copy.h -- row-band parallel host copy used to stage PAGEABLE caller images into pinned memory.
//
// A camera-sized BGR image (1280x886: 3.4 MB) costs ~0.3 ms to copy with one thread -- four times the DMA that follows
// (f1, SURVEY.md section 8f: host ingest).  The pool keeps a few worker threads parked on a condition variable; copy_rows()
// splits the rows into one band per thread (the caller copies a band itself) and returns when all bands are done.
// One pool per handle; the handle's entry points are single-threaded by contract, so copy_rows() is never re-entered.
#pragma once
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace rf {

class HostCopyPool {
   public:
    explicit HostCopyPool(int workers) {
        jobs_.resize(workers < 0 ? 0 : workers);
        for (int i = 0; i < (int)jobs_.size(); i++) threads_.emplace_back([this, i] { run(i); });
    }
    ~HostCopyPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_start_.notify_all();
        for (auto &t : threads_) t.join();
    }
    HostCopyPool(const HostCopyPool &) = delete;
    HostCopyPool &operator=(const HostCopyPool &) = delete;

    // dst: packed rows (row_bytes each); src: rows `src_stride` bytes apart
    void copy_rows(uint8_t *dst, const uint8_t *src, size_t row_bytes, size_t src_stride, int rows) {
        const int parts = (int)jobs_.size() + 1;
        if (parts == 1 || (size_t)rows * row_bytes < (1u << 20) || rows < parts) {   // small images: not worth a wake-up
            band(dst, src, row_bytes, src_stride, 0, rows);
            return;
        }
        const int per = (rows + parts - 1) / parts;
        {
            std::lock_guard<std::mutex> lk(m_);
            for (int i = 0; i < (int)jobs_.size(); i++) {
                const int r0 = std::min(rows, (i + 1) * per), r1 = std::min(rows, (i + 2) * per);
                jobs_[i] = Job{dst, src, row_bytes, src_stride, r0, r1};
            }
            pending_ = (int)jobs_.size();
            gen_++;
        }
        cv_start_.notify_all();
        band(dst, src, row_bytes, src_stride, 0, std::min(rows, per));
        std::unique_lock<std::mutex> lk(m_);
        cv_done_.wait(lk, [this] { return pending_ == 0; });
    }

   private:
    struct Job {
        uint8_t *dst;
        const uint8_t *src;
        size_t row_bytes, src_stride;
        int r0, r1;
    };
    static void band(uint8_t *dst, const uint8_t *src, size_t row_bytes, size_t src_stride, int r0, int r1) {
        if (r1 <= r0) return;
        if (src_stride == row_bytes) {
            std::memcpy(dst + (size_t)r0 * row_bytes, src + (size_t)r0 * src_stride, (size_t)(r1 - r0) * row_bytes);
            return;
        }
        for (int y = r0; y < r1; y++) std::memcpy(dst + (size_t)y * row_bytes, src + (size_t)y * src_stride, row_bytes);
    }
    void run(int i) {
        unsigned long long seen = 0;
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_start_.wait(lk, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
                j = jobs_[i];
            }
            band(j.dst, j.src, j.row_bytes, j.src_stride, j.r0, j.r1);
            {
                std::lock_guard<std::mutex> lk(m_);
                if (--pending_ == 0) cv_done_.notify_one();
            }
        }
    }
    std::vector<std::thread> threads_;
    std::vector<Job> jobs_;
    std::mutex m_;
    std::condition_variable cv_start_, cv_done_;
    unsigned long long gen_ = 0;
    int pending_ = 0;
    bool stop_ = false;
};

}  // namespace rf
