// pm_reader.hpp -- the host reader of the input path: a database file is read by a few
// threads into a ring of page-locked chunk buffers while the copy engine moves the chunks that
// are already full; the file never sits in host memory as a whole.  Replaces the role of
// MapFileChars / BufferedFileChars (char_io.h:150-190) under fasta_io's readers.
//
// Chunk j covers file bytes [j * chunk, (j+1) * chunk).  The consumer takes the chunks in
// order: acquire(j) blocks until chunk j is in its buffer; release(j, ev) gives the buffer
// back once `ev` (recorded behind the copy that read it) has completed.  Buffers are handed
// to the reader threads in chunk order, so a ring of n_threads + 1 buffers keeps every thread
// and the copy engine busy.
#pragma once
#include <cuda_runtime.h>
#include <fcntl.h>
#include <stdint.h>
#include <sys/stat.h>
#include <unistd.h>

#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace pm {

class ChunkReader {
 public:
  ChunkReader() {}
  ~ChunkReader() { close(); }

  // false: the file cannot be opened / the buffers cannot be allocated (errno-style message in err)
  bool open(const std::string &path, uint64_t chunk_bytes, int n_threads, int n_buffers, std::string *err) {
    close();
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) { if (err) *err = "cannot open " + path; return false; }
    struct stat st;
    if (fstat(fd_, &st) != 0) { if (err) *err = "cannot stat " + path; close(); return false; }
    size_ = (uint64_t)st.st_size;
    chunk_ = chunk_bytes;
    n_chunks_ = (size_ + chunk_ - 1) / chunk_;
    if (n_threads < 1) n_threads = 1;
    if (n_buffers < n_threads + 1) n_buffers = n_threads + 1;
    if ((uint64_t)n_buffers > n_chunks_) n_buffers = (int)(n_chunks_ ? n_chunks_ : 1);
    bufs_.assign((size_t)n_buffers, Buf());
    for (Buf &b : bufs_) {
      // write-combined pages are right for a buffer the CPU only fills and the DMA engine reads,
      // but read(2) fills them through the kernel's copy, which is slow on uncached pages: cached
      if (cudaHostAlloc(&b.p, chunk_, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        if (err) *err = "cannot allocate page-locked chunk buffers";
        close();
        return false;
      }
      cudaEventCreateWithFlags(&b.free_ev, cudaEventDisableTiming);
    }
    state_.assign((size_t)n_chunks_, 0);
    next_ = 0;
    stop_ = false;
    for (int t = 0; t < n_threads && (uint64_t)t < n_chunks_; t++) threads_.emplace_back([this] { run(); });
    return true;
  }
  uint64_t size() const { return size_; }
  uint64_t chunk_bytes() const { return chunk_; }
  uint64_t n_chunks() const { return n_chunks_; }

  // chunk j in page-locked memory (blocks until it has been read); nullptr on a read error
  const uint8_t *acquire(uint64_t j, uint64_t *bytes) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return state_[j] != 0; });
    if (state_[j] < 0) return nullptr;
    *bytes = std::min(chunk_, size_ - j * chunk_);
    return (const uint8_t *)bufs_[j % bufs_.size()].p;
  }
  // the buffer of chunk j may be refilled once `ev` has completed (ev recorded by the caller
  // behind the last device operation that reads the buffer)
  void release(uint64_t j, cudaStream_t stream) {
    Buf &b = bufs_[j % bufs_.size()];
    cudaEventRecord(b.free_ev, stream);
    std::lock_guard<std::mutex> lk(mu_);
    b.released = j + 1;
    cv_.notify_all();
  }
  void close() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      cv_.notify_all();
    }
    for (std::thread &t : threads_) t.join();
    threads_.clear();
    for (Buf &b : bufs_) {
      if (b.free_ev) { cudaEventSynchronize(b.free_ev); cudaEventDestroy(b.free_ev); }
      if (b.p) cudaFreeHost(b.p);
    }
    bufs_.clear();
    if (fd_ >= 0) ::close(fd_);
    fd_ = -1;
  }

 private:
  struct Buf {
    void *p = nullptr;
    cudaEvent_t free_ev = nullptr;
    uint64_t released = 0;   // 1 + the last chunk whose consumer has let go of this buffer
  };
  void run() {
    for (;;) {
      uint64_t j;
      {
        std::unique_lock<std::mutex> lk(mu_);
        if (stop_ || next_ >= n_chunks_) return;
        j = next_++;
        // the buffer's previous tenant (chunk j - n_buffers) must have been released
        const uint64_t nb = bufs_.size();
        if (j >= nb) {
          cv_.wait(lk, [&] { return stop_ || bufs_[j % nb].released == j - nb + 1; });
          if (stop_) return;
        }
      }
      Buf &b = bufs_[j % bufs_.size()];
      if (j >= bufs_.size()) cudaEventSynchronize(b.free_ev);   // ... and its copy must be through
      const uint64_t off = j * chunk_, want = std::min(chunk_, size_ - off);
      uint64_t done = 0;
      bool ok = true;
      while (done < want) {
        const ssize_t r = pread(fd_, (char *)b.p + done, (size_t)(want - done), (off_t)(off + done));
        if (r <= 0) { ok = false; break; }
        done += (uint64_t)r;
      }
      std::lock_guard<std::mutex> lk(mu_);
      state_[j] = ok ? 1 : -1;
      cv_.notify_all();
    }
  }

  int fd_ = -1;
  uint64_t size_ = 0, chunk_ = 0, n_chunks_ = 0, next_ = 0;
  bool stop_ = false;
  std::vector<Buf> bufs_;
  std::vector<signed char> state_;   // per chunk: 0 not read yet, 1 in its buffer, -1 read error
  std::vector<std::thread> threads_;
  std::mutex mu_;
  std::condition_variable cv_;
};

}  // namespace pm
