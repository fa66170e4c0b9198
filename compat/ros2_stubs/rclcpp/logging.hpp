// Minimal stand-in for <rclcpp/logging.hpp> (see rclcpp.hpp in this directory).
//
// The macros count messages per severity in thread-local counters so a harness can observe
// "the reference warned" / "the reference reported an error" without parsing text.  When
// tgx_stub::log_state().throw_on_error is set, RCLCPP_ERROR throws tgx_stub::ErrorLogged: the
// reference follows its in-sampler RCLCPP_ERROR calls with exit(1) (Circle.cpp:85-88,
// Line.cpp:76-79), and a harness that runs many trajectories in one process needs to survive that.
#pragma once

#include <cstdio>
#include <stdexcept>

#include "rclcpp/logger.hpp"

namespace tgx_stub {

struct LogState {
    long n_info = 0;
    long n_warn = 0;
    long n_error = 0;
    bool throw_on_error = false;
    bool echo = false;
};

inline LogState& log_state() {
    static thread_local LogState s;
    return s;
}

struct ErrorLogged : std::runtime_error {
    ErrorLogged() : std::runtime_error("RCLCPP_ERROR") {}
};

template <typename... Args>
inline void emit(const char* sev, const rclcpp::Logger& lg, const char* fmt, Args... args) {
    if (!log_state().echo) return;
    std::fprintf(stderr, "[%s] [%s]: ", sev, lg.get_name());
    if constexpr (sizeof...(Args) == 0) std::fputs(fmt, stderr);
    else std::fprintf(stderr, fmt, args...);
    std::fputc('\n', stderr);
}

}  // namespace tgx_stub

#define RCLCPP_INFO(logger, ...)  do { ++tgx_stub::log_state().n_info;  tgx_stub::emit("INFO", (logger), __VA_ARGS__); } while (0)
#define RCLCPP_WARN(logger, ...)  do { ++tgx_stub::log_state().n_warn;  tgx_stub::emit("WARN", (logger), __VA_ARGS__); } while (0)
#define RCLCPP_ERROR(logger, ...) do { ++tgx_stub::log_state().n_error; tgx_stub::emit("ERROR", (logger), __VA_ARGS__); \
                                       if (tgx_stub::log_state().throw_on_error) throw tgx_stub::ErrorLogged(); } while (0)
