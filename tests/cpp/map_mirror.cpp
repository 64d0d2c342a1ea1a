// map_mirror.cpp -- the reference-side binding of the resident map as INTEGRATION.md section 3.2 sketches it, written out
// in full: a mirror object that lives next to ORB-SLAM2's Map, one forwarding line per mutator (MapPoint::AddObservation /
// EraseObservation / SetBadFlag / SetWorldPos, KeyFrame::SetPose, Map::AddKeyFrame -- src/MapPoint.cc:73-168,
// src/KeyFrame.cc:70, src/Map.cc:32), and the per-frame call that replaces src/Tracking.cc:1849-1930.
// Compiled against include/lccrf.h by tests/test_abi.py (the binding must keep compiling when the ABI moves) and run on
// a B200 by tests/test_gpu_dropin.py: a tiny two-keyframe map, one frame, labels printed.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "lccrf.h"

struct LccrfMapMirror {
    lccrf_map *map = nullptr;
    int kp_stride;
    // the changes of the current step
    int kf_first = -1;
    std::vector<float> kf_pose, kf_intr, kf_bounds, kf_kp, pose, xyz;
    std::vector<int> pose_kf, xyz_id, erase_pt, erase_kf, bad_pt, add_pt, add_kf, add_fid, add_seg{0};

    LccrfMapMirror(lccrf_ctx *ctx, int stride) : kp_stride(stride) {
        if (lccrf_map_create(ctx, stride, &map) != LCCRF_OK) {
            std::fprintf(stderr, "%s\n", lccrf_last_error());
            std::exit(1);
        }
    }
    ~LccrfMapMirror() { lccrf_map_destroy(map); }

    // Map::AddKeyFrame(pKF): Tcw rows 0..2, fx fy cx cy, mnMinX mnMaxX mnMinY mnMaxY, mvKeysUn[i].pt
    void AddKeyFrame(int id, const float Tcw[12], const float intr[4], const float bounds[4], const std::vector<float> &keys_xy) {
        if (kf_first < 0) kf_first = id;
        kf_pose.insert(kf_pose.end(), Tcw, Tcw + 12);
        kf_intr.insert(kf_intr.end(), intr, intr + 4);
        kf_bounds.insert(kf_bounds.end(), bounds, bounds + 4);
        std::vector<float> row(2 * (size_t)kp_stride, 0.f);
        for (size_t i = 0; i < keys_xy.size() && i < row.size(); i++) row[i] = keys_xy[i];
        kf_kp.insert(kf_kp.end(), row.begin(), row.end());
        if (add_seg.back() != (int)add_pt.size()) add_seg.push_back((int)add_pt.size());  // its observations: a new segment
    }
    void SetPose(int kf, const float Tcw[12]) {           // KeyFrame::SetPose
        pose_kf.push_back(kf);
        pose.insert(pose.end(), Tcw, Tcw + 12);
    }
    void SetWorldPos(int pt, float x, float y, float z) {  // MapPoint ctor / MapPoint::SetWorldPos
        xyz_id.push_back(pt);
        xyz.insert(xyz.end(), {x, y, z});
    }
    void AddObservation(int pt, int kf, int idx) {         // MapPoint::AddObservation(pKF, idx)
        add_pt.push_back(pt);
        add_kf.push_back(kf);
        add_fid.push_back(idx);
    }
    void EraseObservation(int pt, int kf) {                // MapPoint::EraseObservation(pKF)
        erase_pt.push_back(pt);
        erase_kf.push_back(kf);
    }
    void SetBadFlag(int pt) { bad_pt.push_back(pt); }      // MapPoint::SetBadFlag

    lccrf_map_delta delta() {
        if (add_seg.back() != (int)add_pt.size()) add_seg.push_back((int)add_pt.size());
        lccrf_map_delta d = {};
        d.kf_first = kf_first < 0 ? 0 : kf_first;
        d.kf_count = (int)(kf_pose.size() / 12);
        d.kf_pose = kf_pose.data();
        d.kf_intr = kf_intr.data();
        d.kf_bounds = kf_bounds.data();
        d.kf_keypoints = kf_kp.data();
        d.n_pose = (int)pose_kf.size();
        d.pose_kf = pose_kf.data();
        d.pose = pose.data();
        d.n_xyz = (int)xyz_id.size();
        d.xyz_id = xyz_id.data();
        d.xyz = xyz.data();
        d.n_erase = (int)erase_pt.size();
        d.erase_pt = erase_pt.data();
        d.erase_kf = erase_kf.data();
        d.n_bad = (int)bad_pt.size();
        d.bad_pt = bad_pt.data();
        d.n_add = (int)add_pt.size();
        d.add_pt = add_pt.data();
        d.add_kf = add_kf.data();
        d.add_fid = add_fid.data();
        d.n_add_seg = (int)add_seg.size() - 1;
        d.add_seg_ptr = add_seg.data();
        return d;
    }
    void clear() {
        kf_first = -1;
        for (auto *v : {&kf_pose, &kf_intr, &kf_bounds, &kf_kp, &pose, &xyz}) v->clear();
        for (auto *v : {&pose_kf, &xyz_id, &erase_pt, &erase_kf, &bad_pt, &add_pt, &add_kf, &add_fid}) v->clear();
        add_seg.assign(1, 0);
    }
};

int main() {
    lccrf_ctx *ctx = nullptr;
    if (lccrf_ctx_create(0, &ctx) != LCCRF_OK) {
        std::fprintf(stderr, "%s\n", lccrf_last_error());
        return 2;  // no B200 here: the binding compiled and linked, which is what the CPU test checks
    }
    const int stride = 64, N = 8;
    LccrfMapMirror m(ctx, stride);
    const float I[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}, T2[12] = {1, 0, 0, -0.1f, 0, 1, 0, 0, 0, 0, 1, 0};
    const float intr[4] = {535.4f, 539.2f, 320.1f, 247.6f}, bounds[4] = {0, 640, 0, 480};
    std::vector<float> keys(2 * stride), kp2d(2 * N);
    for (int i = 0; i < N; i++) {  // points on a fronto-parallel plane, keypoints = their projections (+ a drift for the last two)
        const float x = -0.7f + 0.2f * i, y = 0.1f * (i % 3), z = 2.5f;
        m.SetWorldPos(i, x, y, z);
        keys[2 * i] = intr[0] * x / z + intr[2] + (i >= N - 2 ? 12.f : 0.3f);
        keys[2 * i + 1] = intr[1] * y / z + intr[3];
        kp2d[2 * i] = keys[2 * i];
        kp2d[2 * i + 1] = keys[2 * i + 1];
    }
    m.AddKeyFrame(0, I, intr, bounds, keys);
    for (int i = 0; i < N; i++) m.AddObservation(i, 0, i);
    m.AddKeyFrame(1, T2, intr, bounds, keys);
    for (int i = 0; i < N; i += 2) m.AddObservation(i, 1, i);
    lccrf_slam_params prm = {10.f, 30.f, 1.7f, 0.6f, 5.4f, 1.5f, 0.3f, 0.2f, 0.5f, 18.f, 2.75f, 0.8f, 0.7f, 5};
    const int prob_ptr[2] = {0, N};
    const float en[3] = {0.6931472f, 1.2039728f, 0.3566749f};
    lccrf_frames *fr = nullptr;
    std::vector<int> ids(N);
    for (int i = 0; i < N; i++) ids[i] = i;
    std::vector<short> label(N);
    std::vector<float> prob(2 * N);
    lccrf_map_delta d = m.delta();
    int rc = lccrf_frames_create(ctx, 1, prob_ptr, &prm, en, &fr);
    if (rc == LCCRF_OK) rc = lccrf_frames_submit_visible(fr, 0, m.map, &d, ids.data(), kp2d.data(), nullptr, label.data(), prob.data());
    if (rc == LCCRF_OK) rc = lccrf_frames_wait(fr, 0);
    if (rc != LCCRF_OK) {
        std::fprintf(stderr, "%s\n", lccrf_last_error());
        return 1;
    }
    m.clear();
    int n_kf = 0, n_pt = 0;
    long long n_obs = 0;
    lccrf_map_sizes(m.map, &n_kf, &n_pt, &n_obs, nullptr, nullptr);
    std::printf("map: %d keyframes, %d points, %lld observations; labels:", n_kf, n_pt, n_obs);
    for (int i = 0; i < N; i++) std::printf(" %d", (int)label[i]);
    std::printf("\n");
    lccrf_frames_destroy(fr);
    return 0;
}
