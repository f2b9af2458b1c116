// TEST INFRASTRUCTURE. Runs the per-thread body of the position-split exchange kernel
// (panagram_b200/csrc/pk_gather.cuh, the code gather_slice_kernel executes per chunk) on the CPU, chunk by
// chunk, against a straightforward restatement of its contract. Usage: gather_host_check   (exit 0 = all cases equal)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../panagram_b200/csrc/pk_gather.cuh"

static uint64_t rng_state = 88172645463325252ull;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

template <int RMAX, int W> static void run_all(const PkgArgs &a) {
    for (uint64_t t = 0; t < a.n_chunks; t++) pkg_gather_chunk<RMAX, W>(a, t);
}
template <int RMAX> static void dispatch_w(const PkgArgs &a) {
    switch (a.w) {
        case 1: run_all<RMAX, 1>(a); break;
        case 2: run_all<RMAX, 2>(a); break;
        case 4: run_all<RMAX, 4>(a); break;
        case 8: run_all<RMAX, 8>(a); break;
        case 16: run_all<RMAX, 16>(a); break;
        default: run_all<RMAX, 0>(a); break;
    }
}
static void dispatch(const PkgArgs &a, int force_rmax) {          // as pk_launch_gather_slice picks the instantiation
    const uint32_t r = force_rmax ? (uint32_t)force_rmax : a.n_ranks;
    if (r <= 2) dispatch_w<2>(a); else if (r <= 4) dispatch_w<4>(a); else if (r <= 8) dispatch_w<8>(a); else dispatch_w<16>(a);
}

static int one_case(uint32_t R, uint32_t w, uint64_t plane_rows, uint32_t n_segs, uint32_t row_bytes, uint32_t pad, int force_rmax) {
    // planes: 16-byte aligned, plane_rows * w bytes (+ slack the kernel must not need)
    std::vector<std::vector<uint8_t>> store(R);
    PkgArgs a{};
    for (uint32_t q = 0; q < R; q++) {
        store[q].resize(plane_rows * w + 15);
        uint8_t *p = store[q].data() + ((16 - ((uintptr_t)store[q].data() & 15)) & 15);
        for (uint64_t i = 0; i < plane_rows * w; i++) p[i] = (uint8_t)rnd();
        
        a.planes[q] = p;
    }
    // disjoint ascending segments with gaps, output rows back to back
    std::vector<PkgSeg> segs(n_segs);
    uint64_t src = rnd() % 5, dst = 0;
    for (uint32_t s = 0; s < n_segs; s++) {
        const uint64_t left = plane_rows > src ? plane_rows - src : 0;
        uint64_t n = left ? rnd() % (left / (n_segs - s) + 1) : 0;
        if (s == n_segs - 1 && (rnd() & 1)) n = left;                                   // reach the very end of the planes
        segs[s] = PkgSeg{src, n, dst, 0};
        src += n + rnd() % 40; dst += n;
        if (src > plane_rows) src = plane_rows;
    }
    const uint32_t stride = row_bytes + pad;
    std::vector<uint8_t> out((dst + 1) * stride + 16, 0xAB), want(out);
    uint8_t *o = out.data() + ((16 - ((uintptr_t)out.data() & 15)) & 15), *wn = want.data() + (o - out.data());
    a.n_ranks = R; a.w = w; a.plane_rows = plane_rows; a.segs = segs.data(); a.n_segs = n_segs;
    a.row_stride = stride; a.row_bytes = row_bytes; a.rows = o;
    a.n_chunks = pkg_plan_segments(segs.data(), n_segs, w);
    dispatch(a, force_rmax);
    const uint32_t full = R * w < row_bytes ? R * w : row_bytes;
    for (uint32_t s = 0; s < n_segs; s++)
        for (uint64_t i = 0; i < segs[s].n_rows; i++)
            for (uint32_t q = 0; q < R; q++)
                for (uint32_t b = 0; b < w; b++)
                    if (q * w + b < full) wn[(segs[s].dst_row + i) * stride + q * w + b] = a.planes[q][(segs[s].src_row + i) * w + b];
    if (memcmp(out.data(), want.data(), out.size()) != 0) {
        fprintf(stderr, "MISMATCH R=%u w=%u rows=%llu segs=%u row_bytes=%u pad=%u rmax=%d\n", R, w, (unsigned long long)plane_rows, n_segs, row_bytes, pad, force_rmax);
        return 1;
    }
    return 0;
}

// the output-aligned variant for narrow rows (pkg_gather_dst_chunk), as pk_launch_gather_slice_dst drives it
template <int R, int W> static void run_dst(const PkgArgs &a, uint64_t total) {
    for (uint64_t t = 0; t < a.n_chunks; t++) pkg_gather_dst_chunk<R, W>(a, t, total);
}
static int dst_case(uint32_t R, uint32_t w, uint64_t plane_rows, uint32_t n_segs) {
    std::vector<std::vector<uint8_t>> store(R);
    PkgArgs a{};
    for (uint32_t q = 0; q < R; q++) {
        store[q].resize(plane_rows * w + 15);
        uint8_t *p = store[q].data() + ((16 - ((uintptr_t)store[q].data() & 15)) & 15);
        for (uint64_t i = 0; i < plane_rows * w; i++) p[i] = (uint8_t)rnd();
        a.planes[q] = p;
    }
    std::vector<PkgSeg> segs(n_segs);
    uint64_t src = rnd() % 5, dst = 0;
    for (uint32_t s = 0; s < n_segs; s++) {
        const uint64_t left = plane_rows > src ? plane_rows - src : 0;
        uint64_t n = left ? rnd() % (left / (n_segs - s) + 1) : 0;
        if (s == n_segs - 1 && (rnd() & 1)) n = left;
        segs[s] = PkgSeg{src, n, dst, 0};
        src += n + rnd() % 40; dst += n;
        if (src > plane_rows) src = plane_rows;
    }
    const uint32_t rw = R * w;
    std::vector<uint8_t> out(dst * rw + 48, 0xAB), want(out);
    uint8_t *o = out.data() + ((16 - ((uintptr_t)out.data() & 15)) & 15), *wn = want.data() + (o - out.data());
    a.n_ranks = R; a.w = w; a.plane_rows = plane_rows; a.segs = segs.data(); a.n_segs = n_segs;
    a.row_stride = rw; a.row_bytes = rw; a.rows = o;
    uint64_t total = 0;
    if (!pkg_dst_mode_ok(segs.data(), n_segs, R, w, rw, rw, o, &total) || total != dst) { fprintf(stderr, "dst mode refused R=%u w=%u\n", R, w); return 1; }
    const uint32_t nr = 16 / rw;
    a.n_chunks = (total + nr - 1) / nr;
    switch (R * 100 + w) {
        case 201: run_dst<2, 1>(a, total); break;
        case 202: run_dst<2, 2>(a, total); break;
        case 204: run_dst<2, 4>(a, total); break;
        case 401: run_dst<4, 1>(a, total); break;
        case 402: run_dst<4, 2>(a, total); break;
        case 801: run_dst<8, 1>(a, total); break;
        default: return 1;
    }
    for (uint32_t s = 0; s < n_segs; s++)
        for (uint64_t i = 0; i < segs[s].n_rows; i++)
            for (uint32_t q = 0; q < R; q++)
                for (uint32_t b = 0; b < w; b++) wn[(segs[s].dst_row + i) * rw + q * w + b] = a.planes[q][(segs[s].src_row + i) * w + b];
    if (memcmp(out.data(), want.data(), out.size()) != 0) {
        fprintf(stderr, "DST MISMATCH R=%u w=%u rows=%llu segs=%u\n", R, w, (unsigned long long)plane_rows, n_segs);
        return 1;
    }
    return 0;
}

int main() {
    int bad = 0, n = 0;
    const uint32_t dR[] = {2, 2, 2, 4, 4, 8}, dW[] = {1, 2, 4, 1, 2, 1};
    for (int c = 0; c < 6; c++)
        for (int rep = 0; rep < 40; rep++) {
            const uint64_t rows = rep == 0 ? 1 : rep == 1 ? 7 : 50 + rnd() % 3000;
            bad += dst_case(dR[c], dW[c], rows, 1 + rnd() % 6); n++;
        }
    const uint32_t Rs[] = {1, 2, 3, 4, 5, 8, 16}, ws[] = {1, 2, 3, 4, 8, 16};
    for (uint32_t R : Rs)
        for (uint32_t w : ws)
            for (int rep = 0; rep < 6; rep++) {
                const uint64_t rows = rep == 0 ? 1 : rep == 1 ? 15 : 100 + rnd() % 3000;
                const uint32_t n_segs = 1 + rnd() % 6;
                // full rows, rows with padding in the stride, and a narrow last shard (row_bytes < R * w)
                bad += one_case(R, w, rows, n_segs, R * w, 0, 0); n++;
                bad += one_case(R, w, rows, n_segs, R * w, (uint32_t)(rnd() % 5), 0); n++;
                if (w > 1) { bad += one_case(R, w, rows, n_segs, R * w - 1 - (uint32_t)(rnd() % (w - 1)), 0, 0); n++; }
                if (R > 2) { bad += one_case(R, w, rows, n_segs, R * w, 0, 2); n++; }       // more ranks than the instantiation: byte path
            }
    printf("%d cases, %d mismatches\n", n, bad);
    return bad ? 1 : 0;
}
