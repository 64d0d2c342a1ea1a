// snapshot.cu -- CRF-input snapshot / replay file format.  Host code only (no device work).
//
// The reference has no way to replay a sequence through the CRF: its inputs live in a pointer graph
// (MapPoint::mObservations -> KeyFrame -> mvKeysUn) that exists only while ORB-SLAM runs.  A snapshot file holds,
// frame after frame, the flat restatement of exactly what Tracking::DynamicDetectionWithCRF gathers at
// src/Tracking.cc:1849-1870 and what ComputeMapPointErrAndObserv dereferences at :1803-1839:
//   per point   xyz (MapPoint::GetWorldPos), kp2d (mCurrentFrame.mvKeysUn[i].pt), fid (featureMapAssos[i].fid)
//   per observation   keyframe index + the keypoint it was seen at (pKF->mvKeysUn[idx].pt), CSR by point
//   per keyframe   Tcw rows [Rcw|tcw] (KeyFrame::GetPose), fx fy cx cy (KeyFrame.h:157), mnMinX..mnMaxY (:185-188)
// which is the lccrf_frames_set_map_inputs layout, so a file replays through the device path (and through the
// reference headers in tests) without ORB-SLAM.
//
// Layout (little endian; every array padded to 8 bytes):
//   file header  32 B : "LCCRFSNP", u32 version = 1, u32 endian tag 0x01020304, 16 B reserved (0)
//   frame record 64 B : u32 'FRAM', u32 flags (bit 0: obs_kf stored as uint16), i32 N, i32 nKF, i64 nnz, i64 frame_id,
//                       f64 timestamp, u64 payload bytes, u64 FNV-1a-64 of the payload, u64 reserved (0)
//   payload           : xyz f32[N*3] | obs_ptr i32[N+1] | obs_kf u16/i32[nnz] | obs_uv f32[nnz*2] | kf_pose f32[nKF*12] |
//                       kf_intr f32[nKF*4] | kf_bounds f32[nKF*4] | kp2d f32[N*2] | fid i32[N]
// Records are self-describing and appended one per frame, so a writer can run inside the tracking thread and a
// truncated file (crash) still yields every complete frame.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <unistd.h>

#include "common.cuh"

namespace {

using lccrf::fail;

constexpr uint32_t kVersion = 1, kEndian = 0x01020304u, kFrameMagic = 0x4D415246u;  // "FRAM"
const char kFileMagic[8] = {'L', 'C', 'C', 'R', 'F', 'S', 'N', 'P'};

struct FileHeader {
    char magic[8];
    uint32_t version, endian;
    uint64_t reserved[2];
};
static_assert(sizeof(FileHeader) == 32, "file header is 32 bytes");

struct FrameHeader {
    uint32_t magic, flags;
    int32_t N, nKF;
    int64_t nnz, frame_id;
    double timestamp;
    uint64_t payload_bytes, checksum, reserved;
};
static_assert(sizeof(FrameHeader) == 64, "frame header is 64 bytes");

inline uint64_t pad8(uint64_t b) { return (b + 7) & ~(uint64_t)7; }

inline uint64_t fnv1a(uint64_t h, const void *p, size_t n) {
    const unsigned char *c = (const unsigned char *)p;
    for (size_t i = 0; i < n; i++) {
        h ^= c[i];
        h *= 0x100000001b3ULL;
    }
    return h;
}
constexpr uint64_t kFnvInit = 0xcbf29ce484222325ULL;

struct Sections {  // byte offsets of the payload arrays
    uint64_t xyz, obs_ptr, obs_kf, obs_uv, kf_pose, kf_intr, kf_bounds, kp2d, fid, total;
};

Sections sections(int N, long long nnz, int nKF, int kf_bytes) {
    Sections s;
    uint64_t o = 0;
    auto take = [&](uint64_t bytes) {
        const uint64_t at = o;
        o += pad8(bytes);
        return at;
    };
    s.xyz = take((uint64_t)N * 12);
    s.obs_ptr = take((uint64_t)(N + 1) * 4);
    s.obs_kf = take((uint64_t)nnz * kf_bytes);
    s.obs_uv = take((uint64_t)nnz * 8);
    s.kf_pose = take((uint64_t)nKF * 48);
    s.kf_intr = take((uint64_t)nKF * 16);
    s.kf_bounds = take((uint64_t)nKF * 16);
    s.kp2d = take((uint64_t)N * 8);
    s.fid = take((uint64_t)N * 4);
    s.total = o;
    return s;
}

// Walk the frame records behind the file header.  Returns the byte offset behind the last COMPLETE, well-formed record;
// anything after it (a record the writer died in, or bytes that do not parse as a record) sets *damaged.
uint64_t scan_records(FILE *f, uint64_t size, std::vector<FrameHeader> *hdr, std::vector<uint64_t> *at, int *damaged) {
    uint64_t pos = sizeof(FileHeader);
    *damaged = 0;
    while (pos + sizeof(FrameHeader) <= size) {
        FrameHeader h;
        fseek(f, (long)pos, SEEK_SET);
        if (fread(&h, sizeof(h), 1, f) != 1) break;
        if (h.magic != kFrameMagic || h.N < 0 || h.nKF < 0 || h.nnz < 0 ||
            h.payload_bytes != sections(h.N, h.nnz, h.nKF, (h.flags & 1) ? 2 : 4).total)
            break;  // not a record: everything from here on is unusable, the frames before it are intact
        if (pos + sizeof(FrameHeader) + h.payload_bytes > size) break;  // the writer died inside this frame
        if (hdr) hdr->push_back(h);
        if (at) at->push_back(pos + sizeof(FrameHeader));
        pos += sizeof(FrameHeader) + h.payload_bytes;
    }
    if (pos != size) *damaged = 1;
    return pos;
}

}  // namespace

struct lccrf_snapshot_writer {
    FILE *f = nullptr;
    int frames = 0;
    std::vector<unsigned char> buf;
};

struct lccrf_snapshot_reader {
    FILE *f = nullptr;
    std::vector<FrameHeader> hdr;
    std::vector<uint64_t> at;  // file offset of every payload
    std::vector<unsigned char> buf;
    int truncated = 0;
};

extern "C" {

int lccrf_snapshot_writer_open(const char *path, int append, lccrf_snapshot_writer **out) {
    if (!path || !out) return fail(LCCRF_ERR_ARG, "NULL argument");
    *out = nullptr;
    FILE *f = nullptr;
    bool fresh = true;
    if (append) {
        f = fopen(path, "r+b");
        if (f) {
            FileHeader h;
            if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, kFileMagic, 8) != 0 || h.version != kVersion ||
                h.endian != kEndian) {
                fclose(f);
                return fail(LCCRF_ERR_ARG, std::string(path) + ": not a version-1 LCCRFSNP file");
            }
            // continue behind the last complete record: a tail left by a crash (partial record) is cut off, otherwise
            // the new record would land inside the partial one's declared payload and take the file down with it
            fseek(f, 0, SEEK_END);
            const uint64_t size = (uint64_t)ftell(f);
            int damaged = 0;
            const uint64_t end = scan_records(f, size, nullptr, nullptr, &damaged);
            if (damaged) {
                fflush(f);
                if (ftruncate(fileno(f), (off_t)end) != 0) {
                    fclose(f);
                    return fail(LCCRF_ERR_STATE, std::string(path) + ": cannot cut off the incomplete tail");
                }
            }
            fseek(f, (long)end, SEEK_SET);
            fresh = false;
        }
    }
    if (!f) f = fopen(path, "wb");
    if (!f) return fail(LCCRF_ERR_ARG, std::string("cannot open ") + path + " for writing");
    if (fresh) {
        FileHeader h;
        memset(&h, 0, sizeof(h));
        memcpy(h.magic, kFileMagic, 8);
        h.version = kVersion;
        h.endian = kEndian;
        if (fwrite(&h, sizeof(h), 1, f) != 1) {
            fclose(f);
            return fail(LCCRF_ERR_STATE, "write failed");
        }
    }
    lccrf_snapshot_writer *w = new lccrf_snapshot_writer();
    w->f = f;
    *out = w;
    return LCCRF_OK;
}

int lccrf_snapshot_write_frame(lccrf_snapshot_writer *w, long long frame_id, double timestamp, int N, const float *xyz,
                               const int *obs_ptr, const int *obs_kf, const float *obs_uv, int nKF, const float *kf_pose,
                               const float *kf_intr, const float *kf_bounds, const float *kp2d, const int *fid) {
    if (!w || !w->f) return fail(LCCRF_ERR_ARG, "writer is NULL or closed");
    if (N < 0 || nKF < 0) return fail(LCCRF_ERR_ARG, "negative size");
    if (N > 0 && (!xyz || !obs_ptr || !kp2d)) return fail(LCCRF_ERR_ARG, "NULL argument");
    const long long nnz = N > 0 ? obs_ptr[N] : 0;
    if (N > 0 && obs_ptr[0] != 0) return fail(LCCRF_ERR_ARG, "obs_ptr must start at 0");
    for (int i = 0; i < N; i++)
        if (obs_ptr[i + 1] < obs_ptr[i]) return fail(LCCRF_ERR_ARG, "obs_ptr must be non-decreasing");
    if (nnz > 0 && (!obs_kf || !obs_uv || !kf_pose || !kf_intr || !kf_bounds)) return fail(LCCRF_ERR_ARG, "NULL argument");
    for (long long e = 0; e < nnz; e++)
        if (obs_kf[e] < 0 || obs_kf[e] >= nKF) return fail(LCCRF_ERR_ARG, "obs_kf out of range");
    const int kf_bytes = nKF <= 65536 ? 2 : 4;
    const Sections s = sections(N, nnz, nKF, kf_bytes);
    w->buf.assign(s.total, 0);
    unsigned char *p = w->buf.data();
    if (N) {
        memcpy(p + s.xyz, xyz, (size_t)N * 12);
        memcpy(p + s.obs_ptr, obs_ptr, (size_t)(N + 1) * 4);
        memcpy(p + s.kp2d, kp2d, (size_t)N * 8);
        if (fid) memcpy(p + s.fid, fid, (size_t)N * 4);
        else
            for (int i = 0; i < N; i++) ((int32_t *)(p + s.fid))[i] = i;
    } else {
        memset(p + s.obs_ptr, 0, 4);
    }
    if (nnz) {
        if (kf_bytes == 2)
            for (long long e = 0; e < nnz; e++) ((uint16_t *)(p + s.obs_kf))[e] = (uint16_t)obs_kf[e];
        else
            memcpy(p + s.obs_kf, obs_kf, (size_t)nnz * 4);
        memcpy(p + s.obs_uv, obs_uv, (size_t)nnz * 8);
    }
    if (nKF && kf_pose && kf_intr && kf_bounds) {
        memcpy(p + s.kf_pose, kf_pose, (size_t)nKF * 48);
        memcpy(p + s.kf_intr, kf_intr, (size_t)nKF * 16);
        memcpy(p + s.kf_bounds, kf_bounds, (size_t)nKF * 16);
    }
    FrameHeader h;
    memset(&h, 0, sizeof(h));
    h.magic = kFrameMagic;
    h.flags = kf_bytes == 2 ? 1u : 0u;
    h.N = N;
    h.nKF = nKF;
    h.nnz = nnz;
    h.frame_id = frame_id;
    h.timestamp = timestamp;
    h.payload_bytes = s.total;
    h.checksum = fnv1a(kFnvInit, p, s.total);
    if (fwrite(&h, sizeof(h), 1, w->f) != 1 || (s.total && fwrite(p, 1, s.total, w->f) != s.total))
        return fail(LCCRF_ERR_STATE, "write failed (disk full?)");
    fflush(w->f);
    w->frames++;
    return LCCRF_OK;
}

int lccrf_snapshot_writer_close(lccrf_snapshot_writer *w) {
    if (!w) return LCCRF_OK;
    int rc = LCCRF_OK;
    if (w->f && fclose(w->f) != 0) rc = fail(LCCRF_ERR_STATE, "close failed");
    delete w;
    return rc;
}

int lccrf_snapshot_reader_open(const char *path, lccrf_snapshot_reader **out) {
    if (!path || !out) return fail(LCCRF_ERR_ARG, "NULL argument");
    *out = nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) return fail(LCCRF_ERR_ARG, std::string("cannot open ") + path);
    FileHeader fh;
    if (fread(&fh, sizeof(fh), 1, f) != 1 || memcmp(fh.magic, kFileMagic, 8) != 0) {
        fclose(f);
        return fail(LCCRF_ERR_ARG, std::string(path) + ": not an LCCRFSNP file");
    }
    if (fh.version != kVersion || fh.endian != kEndian) {
        fclose(f);
        return fail(LCCRF_ERR_ARG, std::string(path) + ": unsupported version or byte order");
    }
    fseek(f, 0, SEEK_END);
    const uint64_t size = (uint64_t)ftell(f);
    lccrf_snapshot_reader *r = new lccrf_snapshot_reader();
    r->f = f;
    // index the complete records; a damaged tail (crash inside a frame, or bytes that do not parse) ends the index and
    // is reported through lccrf_snapshot_truncated -- the frames in front of it stay readable
    scan_records(f, size, &r->hdr, &r->at, &r->truncated);
    *out = r;
    return LCCRF_OK;
}

void lccrf_snapshot_reader_close(lccrf_snapshot_reader *r) {
    if (!r) return;
    if (r->f) fclose(r->f);
    delete r;
}

int lccrf_snapshot_num_frames(const lccrf_snapshot_reader *r) { return r ? (int)r->hdr.size() : 0; }
int lccrf_snapshot_truncated(const lccrf_snapshot_reader *r) { return r ? r->truncated : 0; }

int lccrf_snapshot_frame_info(const lccrf_snapshot_reader *r, int i, lccrf_snapshot_info *info) {
    if (!r || !info) return fail(LCCRF_ERR_ARG, "NULL argument");
    if (i < 0 || i >= (int)r->hdr.size()) return fail(LCCRF_ERR_ARG, "frame index out of range");
    const FrameHeader &h = r->hdr[i];
    info->N = h.N;
    info->nKF = h.nKF;
    info->nnz = h.nnz;
    info->frame_id = h.frame_id;
    info->timestamp = h.timestamp;
    info->stored_kf_bytes = (h.flags & 1) ? 2 : 4;
    return LCCRF_OK;
}

int lccrf_snapshot_read_frame(lccrf_snapshot_reader *r, int i, float *xyz, int *obs_ptr, void *obs_kf, int obs_kf_bytes,
                              float *obs_uv, float *kf_pose, float *kf_intr, float *kf_bounds, float *kp2d, int *fid) {
    if (!r) return fail(LCCRF_ERR_ARG, "reader is NULL");
    if (i < 0 || i >= (int)r->hdr.size()) return fail(LCCRF_ERR_ARG, "frame index out of range");
    if (obs_kf && obs_kf_bytes != 2 && obs_kf_bytes != 4) return fail(LCCRF_ERR_ARG, "obs_kf_bytes must be 4 (int32) or 2 (uint16)");
    const FrameHeader &h = r->hdr[i];
    if (obs_kf && obs_kf_bytes == 2 && h.nKF > 65536) return fail(LCCRF_ERR_ARG, "uint16 keyframe indices need nKF <= 65536");
    const int stored = (h.flags & 1) ? 2 : 4;
    const Sections s = sections(h.N, h.nnz, h.nKF, stored);
    r->buf.resize(s.total);
    fseek(r->f, (long)r->at[i], SEEK_SET);
    if (s.total && fread(r->buf.data(), 1, s.total, r->f) != s.total) return fail(LCCRF_ERR_STATE, "short read");
    const unsigned char *p = r->buf.data();
    if (fnv1a(kFnvInit, p, s.total) != h.checksum)
        return fail(LCCRF_ERR_STATE, "frame " + std::to_string(i) + ": checksum mismatch (corrupt payload)");
    // structural validation: what the device path would otherwise have to trust
    const int32_t *ptr = (const int32_t *)(p + s.obs_ptr);
    if (ptr[0] != 0 || ptr[h.N] != h.nnz) return fail(LCCRF_ERR_STATE, "frame " + std::to_string(i) + ": obs_ptr does not span [0, nnz]");
    for (int k = 0; k < h.N; k++)
        if (ptr[k + 1] < ptr[k]) return fail(LCCRF_ERR_STATE, "frame " + std::to_string(i) + ": obs_ptr decreases");
    for (long long e = 0; e < h.nnz; e++) {
        const int k = stored == 2 ? (int)((const uint16_t *)(p + s.obs_kf))[e] : ((const int32_t *)(p + s.obs_kf))[e];
        if (k < 0 || k >= h.nKF) return fail(LCCRF_ERR_STATE, "frame " + std::to_string(i) + ": keyframe index out of range");
    }
    if (xyz) memcpy(xyz, p + s.xyz, (size_t)h.N * 12);
    if (obs_ptr) memcpy(obs_ptr, ptr, (size_t)(h.N + 1) * 4);
    if (obs_kf) {
        if (obs_kf_bytes == stored) memcpy(obs_kf, p + s.obs_kf, (size_t)h.nnz * stored);
        else if (obs_kf_bytes == 4)
            for (long long e = 0; e < h.nnz; e++) ((int32_t *)obs_kf)[e] = ((const uint16_t *)(p + s.obs_kf))[e];
        else
            for (long long e = 0; e < h.nnz; e++) ((uint16_t *)obs_kf)[e] = (uint16_t)((const int32_t *)(p + s.obs_kf))[e];
    }
    if (obs_uv) memcpy(obs_uv, p + s.obs_uv, (size_t)h.nnz * 8);
    if (kf_pose) memcpy(kf_pose, p + s.kf_pose, (size_t)h.nKF * 48);
    if (kf_intr) memcpy(kf_intr, p + s.kf_intr, (size_t)h.nKF * 16);
    if (kf_bounds) memcpy(kf_bounds, p + s.kf_bounds, (size_t)h.nKF * 16);
    if (kp2d) memcpy(kp2d, p + s.kp2d, (size_t)h.N * 8);
    if (fid) memcpy(fid, p + s.fid, (size_t)h.N * 4);
    return LCCRF_OK;
}

}  // extern "C"
