// rigid2d::CircleFitting with the reference's public call surface
// (rigid2d/include/rigid2d/circle_fitting.hpp:18-52) over the C ABI in circle_fit_b200.h; what the landmarks node
// (nuslam/src/landmarks.cpp:133-141) and nuslam/tests/circle_tests.cpp call compiles against this unchanged.
#ifndef CIRCLE_FITTING_INCLUDE_GUARD_HPP
#define CIRCLE_FITTING_INCLUDE_GUARD_HPP

#include <cstdint>
#include <cstdio>
#include <vector>

#include "../circle_fit_b200.h"
#include "rigid2d.hpp"

namespace rigid2d {

class CircleFitting {
  public:
    CircleFitting() {}
    CircleFitting(const CircleFitting& o) : point_cluster(o.point_cluster), xy_cluster(o.xy_cluster), r_cluster(o.r_cluster), keep_(o.keep_) {}
    CircleFitting& operator=(const CircleFitting& o) {
        if (this != &o) {
            point_cluster = o.point_cluster;
            xy_cluster = o.xy_cluster;
            r_cluster = o.r_cluster;
            keep_ = o.keep_;
        }
        return *this;  // the device context is per object and rebuilt lazily
    }
    ~CircleFitting() {
        if (ctx_) circles_destroy(ctx_);
    }

    std::vector<rigid2d::Vector2D> approxCirclePositions(std::vector<double> ranges) {
        clusteringRanges(ranges);
        return classifyCircle(circleRegression());
    }

    void clusteringRanges(std::vector<double> ranges) {
        point_cluster.clear();
        xy_cluster.clear();
        r_cluster.clear();
        if (ranges.size() < 2 || !ensure((int)ranges.size())) return;
        const int M = circles_max_clusters();
        int32_t n = 0;
        std::vector<int32_t> segs(4 * (size_t)M);
        std::vector<double> xy(2 * ranges.size());
        if (check(circles_run_f64(ctx_, ranges.data(), 1, nullptr, nullptr), "circles_run_f64")) return;
        if (check(circles_last_clusters(ctx_, 0, &n, segs.data(), nullptr, nullptr, xy.data()), "circles_last_clusters")) return;
        for (int c = 0; c < n; ++c) {
            std::vector<double> pr;
            std::vector<rigid2d::Vector2D> px;
            for (int part = 0; part < 2; ++part)
                for (int k = 0; k < segs[4 * c + 2 * part + 1]; ++k) {
                    const int i = segs[4 * c + 2 * part] + k;
                    pr.push_back(ranges[i]);
                    px.push_back(rigid2d::Vector2D(xy[2 * i], xy[2 * i + 1]));
                }
            point_cluster.push_back(pr);
            xy_cluster.push_back(px);
        }
    }

    std::vector<rigid2d::Vector2D> circleRegression() {
        std::vector<rigid2d::Vector2D> pos;
        keep_.clear();
        if (xy_cluster.empty() || !ensure(n_beams_ > 0 ? n_beams_ : 360)) return pos;
        std::vector<double> flat;
        std::vector<int32_t> sizes;
        for (const auto& c : xy_cluster) {
            sizes.push_back((int32_t)c.size());
            for (const auto& p : c) {
                flat.push_back(p.x);
                flat.push_back(p.y);
            }
        }
        std::vector<double> cxr(4 * sizes.size());
        std::vector<uint8_t> flags(sizes.size());
        if (check(circles_fit_clusters(ctx_, flat.data(), sizes.data(), (int)sizes.size(), cxr.data(), flags.data()),
                  "circles_fit_clusters"))
            return pos;
        for (size_t c = 0; c < sizes.size(); ++c) {
            pos.push_back(rigid2d::Vector2D(cxr[4 * c], cxr[4 * c + 1]));
            r_cluster.push_back(cxr[4 * c + 2]);
            keep_.push_back((flags[c] & 1) != 0);
        }
        return pos;
    }

    std::vector<rigid2d::Vector2D> classifyCircle(std::vector<rigid2d::Vector2D> circle_positions) {
        std::vector<rigid2d::Vector2D> clean;
        for (size_t c = 0; c < circle_positions.size() && c < keep_.size(); ++c)
            if (keep_[c]) clean.push_back(circle_positions[c]);
        return clean;
    }

    std::vector<std::vector<double>> get_point_cluster() { return point_cluster; }
    std::vector<double> get_r_cluster() { return r_cluster; }
    void set_xy_cluster(std::vector<std::vector<rigid2d::Vector2D>> new_xy_cluster) { xy_cluster = new_xy_cluster; }

    int last_status() const { return status_; }

  private:
    std::vector<std::vector<double>> point_cluster;
    std::vector<std::vector<rigid2d::Vector2D>> xy_cluster;
    std::vector<double> r_cluster;
    std::vector<bool> keep_;
    circles_ctx* ctx_ = nullptr;
    int n_beams_ = 0;
    int status_ = 0;

    bool ensure(int n_beams) {
        if (ctx_ && n_beams_ == n_beams) return true;
        if (ctx_) circles_destroy(ctx_);
        ctx_ = nullptr;
        n_beams_ = n_beams;
        return !check(circles_create(1, n_beams, 64, 0, &ctx_), "circles_create");
    }
    bool check(int rc, const char* what) {
        status_ = rc;
        if (rc != 0) std::fprintf(stderr, "[circle_fit_b200] %s failed (%d): %s\n", what, rc, circles_last_error());
        return rc != 0;
    }
};

}  // namespace rigid2d
#endif
