// Minimal stand-in for <rclcpp/logger.hpp> (see rclcpp.hpp in this directory).
#pragma once

#include <string>

namespace rclcpp {

class Logger {
public:
    Logger() = default;
    explicit Logger(std::string name) : name_(std::move(name)) {}
    const char* get_name() const { return name_.c_str(); }
private:
    std::string name_;
};

inline Logger get_logger(const std::string& name) { return Logger(name); }

}  // namespace rclcpp
