// TEST INFRASTRUCTURE (oracle) - never used by the product path.
//
// Minimal stand-in for the Xilinx HLS `hls::stream<T>` of C simulation, enough
// for the reference-style dataflow kernel printed by oracle/dataflow_kernel.py
// (reference: src/soda/codegen/xilinx/hls_kernel.py:116-148 declares the
// streams, :919-939 reads and writes them).  Under C simulation the modules of
// a dataflow region run one after the other and a stream is an unbounded
// queue; that is what this is.  A chunked queue measured faster than a vector
// with a read cursor (a stream is filled with a whole tensor by its producer
// module before its consumer drains it: no re-allocation copies).
#pragma once

#include <cstddef>
#include <cstdlib>
#include <cstdio>
#include <deque>

namespace hls {

template <typename T>
class stream {
 public:
  stream() : name_("") {}
  explicit stream(const char* name) : name_(name) {}
  stream(const stream&) = delete;
  stream& operator=(const stream&) = delete;

  bool empty() const { return queue_.empty(); }
  size_t size() const { return queue_.size(); }
  void write(const T& value) { queue_.push_back(value); }
  T read() {
    if (empty()) {
      // C simulation prints a warning and returns a default value; a kernel
      // printed by this repo never reads an empty stream, so make it loud
      std::fprintf(stderr, "hls::stream '%s': read while empty\n", name_);
      std::abort();
    }
    T value = queue_.front();
    queue_.pop_front();
    return value;
  }
  void operator<<(const T& value) { write(value); }
  void operator>>(T& value) { value = read(); }

 private:
  const char* name_;
  std::deque<T> queue_;
};

}  // namespace hls
