// Minimal stand-in for <rclcpp/rclcpp.hpp>, written for this repo (NOT ROS code).
//
// ROS2 (rclcpp) is not installed in the build image.  This header provides
//   (1) the handful of names the reference's trajectory classes touch (Trajectory.hpp:15,35; Circle.cpp:35,92):
//       rclcpp::Clock, Clock::SharedPtr, Clock::now(), rclcpp::Time, (Time - Time).seconds();
//   (2) an in-process fake of the rclcpp::Node surface the reference NODE uses (TrajectoryGenerator.cpp:44-93,
//       :99-425): declare_parameter / get_parameter against a harness-filled table, create_subscription /
//       create_wall_timer / create_publisher that record their callbacks and messages instead of talking to DDS, so a
//       test harness can construct the unmodified TrajectoryGenerator, deliver messages and fire its timer tick by tick.
// In a real ROS2 workspace this directory is simply left off the include path.
#pragma once

#include <chrono>
#include <cmath>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <typeindex>
#include <utility>
#include <vector>

#include "builtin_interfaces/msg/time.hpp"
#include "rclcpp/logger.hpp"
#include "rclcpp/logging.hpp"

namespace rclcpp {

class Duration {
public:
    explicit Duration(double s) : s_(s) {}
    double seconds() const { return s_; }
private:
    double s_;
};

class Time {
public:
    Time() : s_(0.0) {}
    explicit Time(double s) : s_(s) {}
    double seconds() const { return s_; }
    Duration operator-(const Time& o) const { return Duration(s_ - o.s_); }
    operator builtin_interfaces::msg::Time() const {
        builtin_interfaces::msg::Time t;
        const double f = std::floor(s_);
        t.sec = (int32_t)f;
        t.nanosec = (uint32_t)((s_ - f) * 1e9);
        return t;
    }
private:
    double s_;
};

class Clock {
public:
    using SharedPtr = std::shared_ptr<Clock>;
    Time now() const {
        using namespace std::chrono;
        return Time(duration<double>(steady_clock::now().time_since_epoch()).count());
    }
};

// ---- the fake node -------------------------------------------------------------------------------------------

enum class DurabilityPolicy { Volatile, TransientLocal };
enum class ReliabilityPolicy { BestEffort, Reliable };

class QoS {
public:
    QoS(int depth) : depth_(depth) {}   // implicit: create_subscription(topic, 1, cb) passes a plain depth
    QoS& durability(DurabilityPolicy) { return *this; }
    QoS& reliability(ReliabilityPolicy) { return *this; }
    int depth() const { return depth_; }
private:
    int depth_;
};

// One parameter value of the harness-filled table (the YAML file of a real launch).
struct ParameterValue {
    enum Kind { NONE, DOUBLE, STRING, DOUBLE_ARRAY } kind = NONE;
    double d = 0.0;
    std::string s;
    std::vector<double> v;
};

}  // namespace rclcpp

namespace tgx_stub {

// What a harness sets before constructing a node: parameter overrides (config/default.yaml) and the namespace.
struct NodeEnvironment {
    std::map<std::string, rclcpp::ParameterValue> overrides;
    std::string ns = "/SQ01";
    bool shutdown_requested = false;
};

inline NodeEnvironment& node_environment() {
    static thread_local NodeEnvironment e;
    return e;
}

}  // namespace tgx_stub

namespace rclcpp {

inline void shutdown() { tgx_stub::node_environment().shutdown_requested = true; }
inline bool ok() { return !tgx_stub::node_environment().shutdown_requested; }
template <class Rep, class Period>
inline void sleep_for(const std::chrono::duration<Rep, Period>&) {}   // nothing to wait for in-process

class TimerBase {
public:
    using SharedPtr = std::shared_ptr<TimerBase>;
    double period_s = 0.0;
    std::function<void()> callback;
};

template <class Msg>
class Publisher {
public:
    using SharedPtr = std::shared_ptr<Publisher<Msg>>;
    void publish(const Msg& m) { sent.push_back(m); }
    std::string topic;
    std::vector<Msg> sent;     // everything published so far, oldest first
};

template <class Msg>
class Subscription {
public:
    using SharedPtr = std::shared_ptr<Subscription<Msg>>;
    std::string topic;
    std::function<void(const Msg&)> callback;
};

class Node {
public:
    explicit Node(const std::string& name) : name_(name), logger_(name), clock_(std::make_shared<Clock>()) {}
    virtual ~Node() {}

    Logger get_logger() const { return logger_; }
    const char* get_namespace() const { return ns_.c_str(); }
    Clock::SharedPtr get_clock() const { return clock_; }
    Time now() const { return clock_->now(); }

    // declare_parameter("alt", 0.0), declare_parameter("traj_type", ""), declare_parameter<std::vector<float>>(...)
    void declare_parameter(const std::string& name, double dflt) {
        ParameterValue p;
        p.kind = ParameterValue::DOUBLE;
        p.d = dflt;
        declare(name, p);
    }
    void declare_parameter(const std::string& name, const char* dflt) {
        ParameterValue p;
        p.kind = ParameterValue::STRING;
        p.s = dflt;
        declare(name, p);
    }
    template <class T>
    void declare_parameter(const std::string& name, const T& dflt) {
        ParameterValue p;
        p.kind = ParameterValue::DOUBLE_ARRAY;
        for (auto x : dflt) p.v.push_back((double)x);
        declare(name, p);
    }

    bool get_parameter(const std::string& name, double& out) const {
        auto it = params_.find(name);
        if (it == params_.end() || it->second.kind != ParameterValue::DOUBLE) return false;
        out = it->second.d;
        return true;
    }
    bool get_parameter(const std::string& name, std::string& out) const {
        auto it = params_.find(name);
        if (it == params_.end() || it->second.kind != ParameterValue::STRING) return false;
        out = it->second.s;
        return true;
    }
    bool get_parameter(const std::string& name, std::vector<double>& out) const {
        auto it = params_.find(name);
        if (it == params_.end() || it->second.kind != ParameterValue::DOUBLE_ARRAY) return false;
        out = it->second.v;
        return true;
    }

    template <class Msg, class Callback>
    typename Subscription<Msg>::SharedPtr create_subscription(const std::string& topic, const QoS&, Callback&& cb) {
        auto s = std::make_shared<Subscription<Msg>>();
        s->topic = topic;
        s->callback = std::forward<Callback>(cb);
        subscriptions_.emplace_back(std::type_index(typeid(Msg)), s);
        return s;
    }
    template <class Msg>
    typename Publisher<Msg>::SharedPtr create_publisher(const std::string& topic, const QoS&) {
        auto p = std::make_shared<Publisher<Msg>>();
        p->topic = topic;
        publishers_.emplace_back(std::type_index(typeid(Msg)), p);
        return p;
    }
    template <class Rep, class Period, class Callback>
    TimerBase::SharedPtr create_wall_timer(const std::chrono::duration<Rep, Period>& period, Callback&& cb) {
        auto t = std::make_shared<TimerBase>();
        t->period_s = std::chrono::duration<double>(period).count();
        t->callback = std::forward<Callback>(cb);
        timers_.push_back(t);
        return t;
    }

    // ---- harness side (not part of rclcpp) ---------------------------------------------------------------
    template <class Msg>
    bool tgx_stub_deliver(const Msg& m) {
        for (auto& kv : subscriptions_)
            if (kv.first == std::type_index(typeid(Msg))) {
                std::static_pointer_cast<Subscription<Msg>>(kv.second)->callback(m);
                return true;
            }
        return false;
    }
    template <class Msg>
    Publisher<Msg>* tgx_stub_publisher() {
        for (auto& kv : publishers_)
            if (kv.first == std::type_index(typeid(Msg))) return std::static_pointer_cast<Publisher<Msg>>(kv.second).get();
        return nullptr;
    }
    bool tgx_stub_fire_timer(size_t i = 0) {
        if (i >= timers_.size()) return false;
        timers_[i]->callback();
        return true;
    }
    double tgx_stub_timer_period(size_t i = 0) const { return i < timers_.size() ? timers_[i]->period_s : 0.0; }

private:
    void declare(const std::string& name, const ParameterValue& dflt) {
        auto& ov = tgx_stub::node_environment().overrides;
        auto it = ov.find(name);
        params_[name] = (it != ov.end()) ? it->second : dflt;
    }

    std::string name_;
    std::string ns_ = tgx_stub::node_environment().ns;
    Logger logger_;
    Clock::SharedPtr clock_;
    std::map<std::string, ParameterValue> params_;
    std::vector<std::pair<std::type_index, std::shared_ptr<void>>> subscriptions_;
    std::vector<std::pair<std::type_index, std::shared_ptr<void>>> publishers_;
    std::vector<TimerBase::SharedPtr> timers_;
};

}  // namespace rclcpp
