// ros_lite — a single-process, deterministic stand-in for the slice of the ROS 1 C++ client API that the reference's
// nodes use (nuslam/src/{slam,unknown_data_assoc,landmarks}.cpp, nurtlesim/src/tube_world.cpp).  TEST INFRASTRUCTURE:
// it exists so that those node sources can be compiled UNMODIFIED, where they lie, against either the reference's own
// rigid2d classes or this repository's drop-in facade, and be driven by a scripted message sequence instead of a ROS
// master (tests/ros_lite/harness.cpp).  Subscriptions, timers and publications go through an in-process registry;
// ros::spin() hands control to the harness, which replays the scenario in simulated time.
#pragma once
#include <cstdint>
#include <cstdio>
#include <functional>
#include <map>
#include <ostream>
#include <string>
#include <typeinfo>
#include <vector>

namespace ros {

struct Duration {
    double sec_ = 0.0;
    Duration() {}
    Duration(double s) : sec_(s) {}  // NOLINT: ros::Duration(0.1) and implicit use both occur
    double toSec() const { return sec_; }
};

namespace lite {
struct Registry {
    double now = 0.0;
    std::map<std::string, std::string> params;  // textual; getParam parses
    struct Sub {
        std::string topic;
        const std::type_info* type;
        std::function<void(const void*)> fn;
    };
    struct Tim {
        double period, next;
        std::function<void()> fn;
    };
    std::vector<Sub> subs;
    std::vector<Tim> timers;
    std::ostream* out = nullptr;  // publications are written here, one line each
    std::function<bool(const std::string&)> keep;  // which topics to record
};
Registry& registry();
void run_scenario();  // harness.cpp

template <class M>
void deliver(const std::string& topic, const M& msg) {
    for (auto& s : registry().subs)
        if (s.topic == topic && *s.type == typeid(M)) s.fn(&msg);
}
}  // namespace lite

struct Time {
    double sec_ = 0.0;
    Time() {}
    explicit Time(double s) : sec_(s) {}
    static Time now() { return Time(lite::registry().now); }
    double toSec() const { return sec_; }
};
struct TimerEvent {};
class Timer {};
class Subscriber {};

class Publisher {
  public:
    Publisher() {}
    explicit Publisher(std::string topic) : topic_(std::move(topic)) {}
    template <class M>
    void publish(const M& m) const {
        auto& r = lite::registry();
        if (r.out && (!r.keep || r.keep(topic_))) {
            char t[40];
            std::snprintf(t, sizeof(t), "%.3f", r.now);
            (*r.out) << t << ' ' << topic_ << ' ';
            lite_dump(*r.out, m);  // found by ADL next to the message type
            (*r.out) << '\n';
        }
        lite::deliver(topic_, m);  // in-process subscribers of the same topic, if any
    }

  private:
    std::string topic_;
};

class NodeHandle {
  public:
    bool getParam(const std::string& key, std::string& v) const {
        auto it = lite::registry().params.find(key);
        if (it == lite::registry().params.end()) return false;
        v = it->second;
        return true;
    }
    bool getParam(const std::string& key, double& v) const {
        std::string s;
        if (!getParam(key, s)) return false;
        v = std::stod(s);
        return true;
    }
    bool getParam(const std::string& key, int& v) const {
        std::string s;
        if (!getParam(key, s)) return false;
        v = std::stoi(s);
        return true;
    }
    bool getParam(const std::string& key, std::vector<double>& v) const {  // comma separated
        std::string s;
        if (!getParam(key, s)) return false;
        v.clear();
        size_t p = 0;
        while (p < s.size()) {
            size_t q = s.find(',', p);
            if (q == std::string::npos) q = s.size();
            v.push_back(std::stod(s.substr(p, q - p)));
            p = q + 1;
        }
        return true;
    }
    template <class M, class T>
    Subscriber subscribe(const std::string& topic, uint32_t, void (T::*fp)(const M&), T* obj) {
        lite::registry().subs.push_back({topic, &typeid(M), [obj, fp](const void* m) { (obj->*fp)(*static_cast<const M*>(m)); }});
        return Subscriber();
    }
    template <class M>
    Publisher advertise(const std::string& topic, uint32_t, bool = false) {
        return Publisher(topic);
    }
    template <class T>
    Timer createTimer(Duration period, void (T::*fp)(const TimerEvent&), T* obj) {
        lite::registry().timers.push_back({period.toSec(), period.toSec(), [obj, fp]() { (obj->*fp)(TimerEvent()); }});
        return Timer();
    }
};

inline void init(int&, char**, const std::string&) {}
inline void spin() { lite::run_scenario(); }

}  // namespace ros

#define ROS_INFO(...) \
    do {              \
    } while (0)
#define ROS_ERROR(...)                      \
    do {                                    \
        std::fprintf(stderr, "[ros_lite] "); \
        std::fprintf(stderr, __VA_ARGS__);  \
        std::fprintf(stderr, "\n");         \
    } while (0)
#define ROS_WARN(...) ROS_INFO(__VA_ARGS__)
