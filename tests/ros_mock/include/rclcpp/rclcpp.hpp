#pragma once
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <typeindex>
#include <unordered_map>
#include <vector>
#include "builtin_interfaces/msg/time.hpp"
namespace rclcpp {
class Time {
 public:
  Time() = default;
  Time(int32_t s, uint32_t ns) : sec_(s), nsec_(ns) {}
  operator builtin_interfaces::msg::Time() const { builtin_interfaces::msg::Time t; t.sec = sec_; t.nanosec = nsec_; return t; }
  double seconds() const { return sec_ + 1e-9 * nsec_; }
 private:
  int32_t sec_ = 0; uint32_t nsec_ = 0;
};
class Duration {
 public:
  Duration(int32_t s, uint32_t ns) : sec_(s), nsec_(ns) {}
  template <class R, class P> Duration(std::chrono::duration<R, P> d) { const auto n = std::chrono::duration_cast<std::chrono::nanoseconds>(d).count(); sec_ = (int32_t)(n / 1000000000); nsec_ = (uint32_t)(n % 1000000000); }
  static Duration from_seconds(double s) { return Duration((int32_t)s, (uint32_t)((s - (int32_t)s) * 1e9)); }
  operator builtin_interfaces::msg::Duration() const { builtin_interfaces::msg::Duration d; d.sec = sec_; d.nanosec = nsec_; return d; }
 private:
  int32_t sec_ = 0; uint32_t nsec_ = 0;
};
class Clock { public: Time now() const { return Time(); } using SharedPtr = std::shared_ptr<Clock>; };
struct Logger {};
struct QoS { QoS(size_t depth = 10) : depth(depth) {} size_t depth; };
struct SubscriptionBase { virtual ~SubscriptionBase() = default; };
template <class Msg> struct Subscription : SubscriptionBase {
  using SharedPtr = std::shared_ptr<Subscription<Msg>>;
  std::function<void(std::shared_ptr<Msg>)> callback;
};
struct PublisherBase { virtual ~PublisherBase() = default; };
template <class Msg> struct Publisher : PublisherBase {
  using SharedPtr = std::shared_ptr<Publisher<Msg>>;
  std::vector<Msg> sent;  // everything published so far (the test harness reads it)
  void publish(const Msg &m) { sent.push_back(m); }
  void publish(std::unique_ptr<Msg> m) { sent.push_back(*m); }
};
struct ParameterValue { double d = 0; long i = 0; bool b = false; std::string s; };
class Node : public std::enable_shared_from_this<Node> {
 public:
  using SharedPtr = std::shared_ptr<Node>;
  explicit Node(const std::string &name) : name_(name) {}
  virtual ~Node() = default;
  // parameters: a test presets overrides in `overrides()` before constructing the node
  static std::map<std::string, ParameterValue> &overrides() { static std::map<std::string, ParameterValue> o; return o; }
  template <class T> T declare_parameter(const std::string &n, const T &dflt) {
    ParameterValue v; store(v, dflt);
    auto it = overrides().find(n);
    params_[n] = it == overrides().end() ? v : it->second;
    T out; load(params_[n], out); return out;
  }
  template <class T> bool get_parameter(const std::string &n, T &out) const {
    auto it = params_.find(n); if (it == params_.end()) return false; load(it->second, out); return true;
  }
  template <class Msg, class CB> typename Subscription<Msg>::SharedPtr create_subscription(const std::string &topic, const QoS &, CB &&cb) {
    auto s = std::make_shared<Subscription<Msg>>(); s->callback = std::forward<CB>(cb); subs_[topic] = s; return s;
  }
  template <class Msg> typename Publisher<Msg>::SharedPtr create_publisher(const std::string &topic, const QoS &) {
    auto p = std::make_shared<Publisher<Msg>>(); pubs_[topic] = p; return p;
  }
  Logger get_logger() const { return Logger(); }
  Time now() const { return Time(); }
  Clock::SharedPtr get_clock() const { return std::make_shared<Clock>(); }
  const char *get_name() const { return name_.c_str(); }
  // harness side
  template <class Msg> void deliver(const std::string &topic, const Msg &m) {
    auto s = std::dynamic_pointer_cast<Subscription<Msg>>(subs_.at(topic)); s->callback(std::make_shared<Msg>(m));
  }
  template <class Msg> std::vector<Msg> &published(const std::string &topic) {
    return std::dynamic_pointer_cast<Publisher<Msg>>(pubs_.at(topic))->sent;
  }
 private:
  static void store(ParameterValue &v, double x) { v.d = x; v.i = (long)x; }
  static void store(ParameterValue &v, int x) { v.i = x; v.d = x; }
  static void store(ParameterValue &v, long x) { v.i = x; v.d = (double)x; }
  static void store(ParameterValue &v, bool x) { v.b = x; }
  static void store(ParameterValue &v, const std::string &x) { v.s = x; }
  static void store(ParameterValue &v, const char *x) { v.s = x; }
  static void load(const ParameterValue &v, double &x) { x = v.d; }
  static void load(const ParameterValue &v, float &x) { x = (float)v.d; }
  static void load(const ParameterValue &v, int &x) { x = (int)v.i; }
  static void load(const ParameterValue &v, long &x) { x = v.i; }
  static void load(const ParameterValue &v, bool &x) { x = v.b; }
  static void load(const ParameterValue &v, std::string &x) { x = v.s; }
  std::string name_;
  std::map<std::string, ParameterValue> params_;
  std::map<std::string, std::shared_ptr<SubscriptionBase>> subs_;
  std::map<std::string, std::shared_ptr<PublisherBase>> pubs_;
};
inline void init(int, char **) {}
inline void init(int, char const *const *) {}
inline void spin(std::shared_ptr<Node>) {}
inline void shutdown() {}
inline bool ok() { return true; }
}  // namespace rclcpp
#ifndef HPP_ROS_MOCK_VERBOSE
#define RCLCPP_INFO(logger, ...) do { (void)(logger); } while (0)
#define RCLCPP_WARN(logger, ...) do { (void)(logger); } while (0)
#else
#define RCLCPP_INFO(logger, ...) do { (void)(logger); std::printf(__VA_ARGS__); std::printf("\n"); } while (0)
#define RCLCPP_WARN(logger, ...) do { (void)(logger); std::printf(__VA_ARGS__); std::printf("\n"); } while (0)
#endif
#define RCLCPP_ERROR(logger, ...) do { (void)(logger); std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#define RCLCPP_DEBUG(logger, ...) do { (void)(logger); } while (0)
