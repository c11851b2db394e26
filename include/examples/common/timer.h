/*
 * examples/common/timer.h -- drop-in for the reference's GJK::Common::PerformanceTimer
 * (reference examples/common/timer.h:15-99): same method names and double-start/stop exceptions.  The GPU timer is
 * implemented on std::chrono around synchronous library calls, so this header needs no CUDA toolkit; the library
 * calls it brackets end with a device synchronisation, exactly as in the reference (openGJK.cu:2983, 3003).
 */
#pragma once
#include <chrono>
#include <stdexcept>

namespace GJK {
namespace Common {

class PerformanceTimer {
 public:
  PerformanceTimer() = default;
  PerformanceTimer(const PerformanceTimer&) = delete;
  PerformanceTimer& operator=(const PerformanceTimer&) = delete;

  void startCpuTimer() { start(cpu_); }
  void endCpuTimer() { stop(cpu_); }
  void startGpuTimer() { start(gpu_); }
  void endGpuTimer() { stop(gpu_); }
  float getCpuElapsedTimeForPreviousOperation() { return cpu_.last_ms; }
  float getGpuElapsedTimeForPreviousOperation() { return gpu_.last_ms; }

 private:
  struct Clock {
    std::chrono::high_resolution_clock::time_point t0;
    bool running = false;
    float last_ms = 0.f;
  };
  static void start(Clock& c) {
    if (c.running) throw std::runtime_error("timer already started");
    c.running = true;
    c.t0 = std::chrono::high_resolution_clock::now();
  }
  static void stop(Clock& c) {
    const auto t1 = std::chrono::high_resolution_clock::now();
    if (!c.running) throw std::runtime_error("timer not started");
    c.last_ms = std::chrono::duration<float, std::milli>(t1 - c.t0).count();
    c.running = false;
  }
  Clock cpu_, gpu_;
};

}  // namespace Common
}  // namespace GJK
