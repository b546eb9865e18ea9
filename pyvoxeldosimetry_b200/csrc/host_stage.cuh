// Host-buffer staging engine of the end-to-end path (pvd_stager_*, pvd_stage_h2d, pvd_stage_d2h).
//
// The reference's calling convention is host NumPy arrays in, a host NumPy array out (core/kernel_convolution.py:48-76);
// the arrays are pageable and usually float64 (np.zeros in every example).  A plain cudaMemcpy from pageable memory runs
// through the driver's single staging thread (measured on the B200 box: 37.6 ms for one 419 MB volume, against 7.6 ms
// for the same bytes from pinned memory), and a float64 volume would cross the link at twice the bytes the engine needs.
// This engine keeps a ring of pinned chunks and a pool of host threads:
//   H2D  worker threads claim chunks in order, copy (float32 / int16 / uint16) or convert (float64 -> float32) their
//        piece of the caller's array into a free ring slot and enqueue that slot's cudaMemcpyAsync on the caller's
//        stream; a slot is reused only after the event recorded behind its previous copy has completed.  The call
//        returns when the caller's array has been read completely (it may be reused) and every copy is enqueued.
//   D2H  the calling thread enqueues chunk copies into free ring slots; workers wait for a slot's event and copy
//        (float32) or widen (float64) it into the caller's array; the call returns when the array is complete.
// Eight threads moved 71 GB/s pageable -> pinned on the box (16: 48 GB/s, profiles/r02_link_peak_1gpu.json), above the
// 55 GB/s of the PCIe link, so the link stays the limit.  While a download runs beside an upload the HOST memory system is
// what saturates (measured: staged upload + pinned download of one volume each 14.2 ms against 8.6 ms pinned both ways,
// profiles/r02_e2e_breakdown.jsonl), so the staging copies use non-temporal stores: no read-for-ownership of the
// destination lines, a third less host-memory traffic per staged byte.
#pragma once
#include "pvd_common.cuh"

#ifndef PVD_EMULATE
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include "host_copy.h"

namespace pvd {

enum HostDtype { HD_F32 = 0, HD_F64 = 1, HD_I16 = 2, HD_U16 = 3 };
inline size_t host_elem_bytes(int dt) { return dt == HD_F64 ? 8 : (dt == HD_F32 ? 4 : 2); }
inline size_t dev_elem_bytes(int dt) { return (dt == HD_F64 || dt == HD_F32) ? 4 : 2; }  // float64 is narrowed on the host

class Stager {
public:
    Stager(int threads, size_t chunk_bytes, int ring) : nthreads_(threads), chunk_(chunk_bytes), ring_(ring) {}
    ~Stager() { shutdown(); }

    cudaError_t init() {
        cudaError_t e = cudaHostAlloc(&pinned_, chunk_ * ring_, cudaHostAllocPortable);
        if (e != cudaSuccess) return e;
        ev_.resize(ring_);
        for (int i = 0; i < ring_; ++i) {
            e = cudaEventCreateWithFlags(&ev_[i], cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
        ev_used_.assign(ring_, 0);
        cudaGetDevice(&device_);
        for (int t = 0; t < nthreads_; ++t) workers_.emplace_back([this] { worker_loop(); });
        return cudaSuccess;
    }

    // ---- host -> device
    cudaError_t h2d(const void* src, int dtype, void* d_dst, size_t n, cudaStream_t stream) {
        if (n == 0) return cudaSuccess;
        const size_t deb = dev_elem_bytes(dtype), heb = host_elem_bytes(dtype);
        const size_t per = chunk_ / deb;  // elements per chunk
        const size_t nchunks = (n + per - 1) / per;
        std::lock_guard<std::mutex> call(call_);  // one transfer at a time per stager
        err_.store(cudaSuccess);
        next_.store(0);
        reset_turns();
        run([&] {
            cudaSetDevice(device_);
            for (;;) {
                const size_t c = next_.fetch_add(1);
                if (c >= nchunks) break;
                const int slot = (int)(c % ring_);
                // slots are claimed in order and every claim of a slot follows the previous claim of the same slot by a
                // whole ring, so waiting on the slot's event orders reuse; the event is recorded by the previous owner
                wait_slot_turn(slot, c / ring_);
                char* stage = pinned_ + (size_t)slot * chunk_;
                const size_t e0 = c * per, cnt = (e0 + per <= n) ? per : n - e0;
                if (ev_used_[slot]) {
                    cudaError_t e = cudaEventSynchronize(ev_[slot]);
                    if (e != cudaSuccess) err_.store(e);
                }
                if (dtype == HD_F64) {
                    stream_narrow(reinterpret_cast<float*>(stage), static_cast<const double*>(src) + e0, cnt);
                } else {
                    stream_copy(stage, static_cast<const char*>(src) + e0 * heb, cnt * heb);
                }
                {
                    // one enqueue at a time keeps (copy, event) pairs adjacent in the stream
                    std::lock_guard<std::mutex> lk(enq_);
                    cudaError_t e = cudaMemcpyAsync(static_cast<char*>(d_dst) + e0 * deb, stage, cnt * deb, cudaMemcpyHostToDevice, stream);
                    if (e == cudaSuccess) e = cudaEventRecord(ev_[slot], stream);
                    if (e != cudaSuccess) err_.store(e);
                    ev_used_[slot] = 1;
                }
                finish_slot_turn(slot);
            }
        });
        return (cudaError_t)err_.load();
    }

    // ---- device (float32) -> host (float32 or float64)
    cudaError_t d2h(const float* d_src, void* dst, int dtype, size_t n, cudaStream_t stream) {
        if (n == 0) return cudaSuccess;
        const size_t per = chunk_ / 4;
        const size_t nchunks = (n + per - 1) / per;
        std::lock_guard<std::mutex> call(call_);
        err_.store(cudaSuccess);
        next_.store(0);
        issued_.store(0);
        reset_turns();
        // producer: this thread enqueues the copies as ring slots become free; consumers: the workers
        std::thread producer([&] {
            cudaSetDevice(device_);
            for (size_t c = 0; c < nchunks; ++c) {
                const int slot = (int)(c % ring_);
                wait_slot_turn(slot, c / ring_);  // the worker that drained the slot's previous chunk has finished
                const size_t e0 = c * per, cnt = (e0 + per <= n) ? per : n - e0;
                cudaError_t e = cudaSuccess;
                if (c < (size_t)ring_ && ev_used_[slot]) e = cudaEventSynchronize(ev_[slot]);  // an earlier transfer (any stream) is done with the slot
                if (e == cudaSuccess)
                    e = cudaMemcpyAsync(pinned_ + (size_t)slot * chunk_, d_src + e0, cnt * 4, cudaMemcpyDeviceToHost, stream);
                if (e == cudaSuccess) e = cudaEventRecord(ev_[slot], stream);
                if (e != cudaSuccess) err_.store(e);
                ev_used_[slot] = 1;
                {
                    std::lock_guard<std::mutex> lk(m_);
                    issued_.store(c + 1);
                }
                cv_.notify_all();
            }
        });
        run([&] {
            cudaSetDevice(device_);
            for (;;) {
                const size_t c = next_.fetch_add(1);
                if (c >= nchunks) break;
                {
                    std::unique_lock<std::mutex> lk(m_);
                    cv_.wait(lk, [&] { return issued_.load() > c; });
                }
                const int slot = (int)(c % ring_);
                cudaError_t e = cudaEventSynchronize(ev_[slot]);
                if (e != cudaSuccess) err_.store(e);
                const float* s = reinterpret_cast<const float*>(pinned_ + (size_t)slot * chunk_);
                const size_t e0 = c * per, cnt = (e0 + per <= n) ? per : n - e0;
                if (dtype == HD_F64) {
                    stream_widen(static_cast<double*>(dst) + e0, s, cnt);
                } else {
                    stream_copy(static_cast<float*>(dst) + e0, s, cnt * 4);
                }
                finish_slot_turn(slot);
            }
        });
        producer.join();
        return (cudaError_t)err_.load();
    }

    int threads() const { return nthreads_; }

private:
    // Per-slot turn counter: chunk c of a transfer may touch slot c % ring only after chunk c - ring has released it.
    void wait_slot_turn(int slot, size_t turn) {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return turn_[slot] == turn; });
    }
    void finish_slot_turn(int slot) {
        {
            std::lock_guard<std::mutex> lk(m_);
            ++turn_[slot];
        }
        cv_.notify_all();
    }
    void reset_turns() {
        std::lock_guard<std::mutex> lk(m_);
        turn_.assign(ring_, 0);
    }
    // run `job` on every worker and wait for all of them
    void run(const std::function<void()>& job) {
        {
            std::lock_guard<std::mutex> lk(m_);
            job_ = &job;
            pending_ = nthreads_;
            ++generation_;
        }
        cv_.notify_all();
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return pending_ == 0; });
        job_ = nullptr;
    }
    void worker_loop() {
        unsigned long long seen = 0;
        for (;;) {
            const std::function<void()>* job;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
                job = job_;
            }
            (*job)();
            {
                std::lock_guard<std::mutex> lk(m_);
                --pending_;
            }
            cv_.notify_all();
        }
    }
    void shutdown() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& w : workers_)
            if (w.joinable()) w.join();
        workers_.clear();
        for (auto& e : ev_) cudaEventDestroy(e);
        ev_.clear();
        if (pinned_) cudaFreeHost(pinned_);
        pinned_ = nullptr;
    }

    int nthreads_;
    size_t chunk_;
    int ring_;
    int device_ = 0;
    char* pinned_ = nullptr;
    std::vector<cudaEvent_t> ev_;
    std::vector<char> ev_used_;
    std::vector<std::thread> workers_;
    std::mutex m_, enq_, call_;
    std::condition_variable cv_;
    std::vector<size_t> turn_;
    const std::function<void()>* job_ = nullptr;
    int pending_ = 0;
    unsigned long long generation_ = 0;
    bool stop_ = false;
    std::atomic<size_t> next_{0}, issued_{0};
    std::atomic<int> err_{0};
};

}  // namespace pvd
#endif  // !PVD_EMULATE
