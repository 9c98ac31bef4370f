// ros_lite harness: replays a scripted scenario against a reference node compiled with -Dmain=node_main
// (see ros/ros.h in this directory).  TEST INFRASTRUCTURE.
//
//   ROS_LITE_SCENARIO  text file:   P <key> <value>                       parameter server entry
//                                   E <t> J <lw> <rw> <dl> <dr>           sensor_msgs/JointState on joint_states
//                                   E <t> F <n> {<x> <y> <action>} x n    visualization_msgs/MarkerArray on fake_sensor
//                                   E <t> C <n> {<x> <y>} x n             visualization_msgs/MarkerArray on scan_sensor
//                                   E <t> S <n> {<range>} x n             sensor_msgs/LaserScan on scan
//                                   E <t> V <v> <w>                       geometry_msgs/Twist on cmd_vel
//                                   END <t>
//   ROS_LITE_OUT       every publication, one line: "<t> <topic> <payload>" (doubles with 17 digits)
//   ROS_LITE_TOPICS    comma-separated topics to record (default: all)
// Simulated time advances in 10 ms ticks; within a tick the due events are delivered first, then the due timers fire
// in the order the node created them.
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

#include "geometry_msgs/Twist.h"
#include "ros/ros.h"
#include "sensor_msgs/JointState.h"
#include "sensor_msgs/LaserScan.h"
#include "visualization_msgs/MarkerArray.h"

int node_main(int argc, char** argv);

namespace {
struct Event {
    double t;
    std::string body;
};
std::vector<Event> g_events;
double g_end = 0.0;
std::ofstream g_out;
std::vector<std::string> g_topics;

struct Loader {
    Loader() {
        const char* path = std::getenv("ROS_LITE_SCENARIO");
        if (!path) return;
        std::ifstream in(path);
        std::string line;
        while (std::getline(in, line)) {
            std::istringstream ss(line);
            std::string kind;
            ss >> kind;
            if (kind == "P") {
                std::string k, v;
                ss >> k;
                std::getline(ss, v);
                const size_t a = v.find_first_not_of(' ');
                ros::lite::registry().params[k] = a == std::string::npos ? "" : v.substr(a);
            } else if (kind == "E") {
                Event e;
                ss >> e.t;
                std::getline(ss, e.body);
                g_events.push_back(e);
            } else if (kind == "END") {
                ss >> g_end;
            }
        }
        if (const char* o = std::getenv("ROS_LITE_OUT")) {
            g_out.open(o);
            ros::lite::registry().out = &g_out;
        }
        if (const char* t = std::getenv("ROS_LITE_TOPICS")) {
            std::istringstream ts(t);
            std::string tok;
            while (std::getline(ts, tok, ',')) g_topics.push_back(tok);
            ros::lite::registry().keep = [](const std::string& topic) {
                for (const auto& k : g_topics)
                    if (k == topic) return true;
                return false;
            };
        }
    }
};
}  // namespace

namespace ros {
namespace lite {
Registry& registry() {
    static Registry r;
    return r;
}

static void dispatch(const std::string& body) {
    std::istringstream ss(body);
    std::string kind;
    ss >> kind;
    if (kind == "J") {
        sensor_msgs::JointState j;
        double v[4];
        ss >> v[0] >> v[1] >> v[2] >> v[3];
        j.position = {v[0], v[1]};
        j.velocity = {v[2], v[3]};
        deliver("joint_states", j);
    } else if (kind == "F" || kind == "C") {
        visualization_msgs::MarkerArray a;
        int n = 0;
        ss >> n;
        for (int i = 0; i < n; ++i) {
            visualization_msgs::Marker m;
            m.id = i;
            ss >> m.pose.position.x >> m.pose.position.y;
            if (kind == "F") ss >> m.action;
            a.markers.push_back(m);
        }
        deliver(kind == "F" ? "fake_sensor" : "scan_sensor", a);
    } else if (kind == "S") {
        sensor_msgs::LaserScan s;
        int n = 0;
        ss >> n;
        s.ranges.resize(n);
        for (int i = 0; i < n; ++i) ss >> s.ranges[i];
        deliver("scan", s);
    } else if (kind == "V") {
        geometry_msgs::Twist t;
        ss >> t.linear.x >> t.angular.z;
        deliver("cmd_vel", t);
    }
}

void run_scenario() {
    Registry& r = registry();
    size_t next = 0;
    for (long k = 1; k * 0.01 <= g_end + 1e-9; ++k) {
        r.now = k / 100.0;
        while (next < g_events.size() && g_events[next].t <= r.now + 1e-9) dispatch(g_events[next++].body);
        for (auto& t : r.timers)
            if (t.next <= r.now + 1e-9) {
                t.next += t.period;
                t.fn();
            }
    }
    if (g_out.is_open()) g_out.flush();
}
}  // namespace lite
}  // namespace ros

static Loader g_loader;

int main(int argc, char** argv) { return node_main(argc, argv); }
