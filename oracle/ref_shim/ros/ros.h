// TEST INFRASTRUCTURE shim: just enough of roscpp for img_env.cpp to compile
// unmodified without a ROS master. Nothing is published or served.
#pragma once
#include <cstdio>
#include <string>
#include <std_msgs/Header.h>
#define ROS_INFO(...) do { } while (0)
namespace ros {
struct Publisher {
    std::string topic;
    std::string getTopic() const { return topic; }
    template <class M> void publish(const M&) const {}
};
struct ServiceServer {};
struct NodeHandle {
    NodeHandle() {}
    explicit NodeHandle(const std::string&) {}
    template <class M> Publisher advertise(const std::string& t, int) { Publisher p; p.topic = t; return p; }
    template <class T, class Req, class Res> ServiceServer advertiseService(const std::string&, bool (T::*)(Req&, Res&), T*) { return ServiceServer(); }
    template <class T> void param(const std::string&, T& v, const T& d) { v = d; }
};
inline void init(int&, char**, const std::string&) {}
inline void spin() {}
}
