// rt_matcher.cpp -- order-preserving batched cross-device signal matcher behind include/rt_matcher.h (host code).
//
// Restates radiotracking/match.py:54-82 (SignalMatcher.add / consume) and the MatchingSignal properties and
// predicates of radiotracking/__init__.py:285-406 on integer microseconds and float64, one native loop per batch.
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/rt_engine.h"
#include "../../include/rt_matcher.h"

int rt_internal_fail(int code, const char* msg);   // rt_engine.cu: sets the thread-local rt_last_error() text

namespace {

struct Group {
    std::vector<rt_match_signal> members;   // dict order of MatchingSignal._sigs: first insertion per device
    int64_t ts = 0, duration = 0;           // min member ts, max member duration (__init__.py:298-317)
    double frequency = 0;                   // statistics.median of the member frequencies (__init__.py:319-329)

    void refresh() {
        ts = members[0].ts_us;
        duration = members[0].duration_us;
        double f[8];
        std::vector<double> big;
        const size_t n = members.size();
        double* fr = f;
        if (n > 8) { big.resize(n); fr = big.data(); }
        for (size_t i = 0; i < n; ++i) {
            ts = std::min(ts, members[i].ts_us);
            duration = std::max(duration, members[i].duration_us);
            fr[i] = members[i].frequency;
        }
        std::sort(fr, fr + n);
        // statistics.median: middle element, or (data[i - 1] + data[i]) / 2 for an even count
        frequency = (n & 1) ? fr[n / 2] : (fr[n / 2 - 1] + fr[n / 2]) / 2;
    }
};

}  // namespace

// what the first-fit walk reads of a group: 24 contiguous bytes instead of a pointer chase per group
struct Hot {
    int64_t ts, duration;
    double frequency;
};

struct rt_matcher {
    int64_t timeout = 0, time_diff = 0, half_duration_diff = 0;
    bool use_duration = false;
    double half_bandwidth = 0;
    std::vector<Group> open;                 // SignalMatcher._matched, creation order
    std::vector<Hot> hot;                    // open[i]'s derived views; duration < 0 marks a closed group awaiting compaction
    size_t head = 0, n_dead = 0;             // entries before `head` are closed; closed entries in [head, size)
    std::vector<int64_t> out_sizes, out_ids; // emitted groups not drained yet
};

namespace {

// MatchingSignal.has_member (__init__.py:337-383).  The reference returns at the first failing comparison; the
// comparisons have no side effects, so evaluating all of them and combining the results is the same predicate --
// and branch-free: which side of a group's frequency a signal falls on is a coin toss for the branch predictor.
inline bool has_member(const rt_matcher& m, const Hot& g, const rt_match_signal& s) {
    unsigned ok = (unsigned)!(s.frequency - m.half_bandwidth > g.frequency) & (unsigned)!(s.frequency + m.half_bandwidth < g.frequency) &
                  (unsigned)!(s.ts_us - m.time_diff > g.ts + g.duration) & (unsigned)!((s.ts_us + s.duration_us) + m.time_diff < g.ts);
    if (m.use_duration)
        ok &= (unsigned)!(s.duration_us - m.half_duration_diff > g.duration) & (unsigned)!(s.duration_us + m.half_duration_diff < g.duration);
    return ok != 0;
}

// MatchingSignal.add_member (__init__.py:385-406)
inline void add_member(Group& g, const rt_match_signal& s) {
    for (auto& mem : g.members)
        if (mem.device == s.device) {
            if (mem.avg < s.avg) { mem = s; g.refresh(); }      // louder detection of the same device replaces, in place
            return;
        }
    g.members.push_back(s);
    g.refresh();
}

inline void emit(rt_matcher& m, const Group& g) {
    m.out_sizes.push_back((int64_t)g.members.size());
    for (const auto& mem : g.members) m.out_ids.push_back(mem.id);
}

}  // namespace

extern "C" {

int rt_matcher_create(int64_t timeout_us, int64_t time_diff_us, double bandwidth_hz, int64_t duration_diff_us, rt_matcher** out) {
    if (!out) return rt_internal_fail(RT_ERR_INVALID, "null argument");
    *out = nullptr;
    if (!(bandwidth_hz == bandwidth_hz)) return rt_internal_fail(RT_ERR_INVALID, "bandwidth is NaN");
    rt_matcher* m = new rt_matcher();
    m->timeout = timeout_us;
    m->time_diff = time_diff_us;
    m->half_bandwidth = bandwidth_hz / 2;                   // `bandwidth / 2` of __init__.py:363,366
    m->use_duration = duration_diff_us > 0;                 // `if duration_diff:` (timedelta truthiness)
    if (m->use_duration) {
        // timedelta / 2 rounds half-microseconds to even (CPython timedelta.__truediv__)
        int64_t q = duration_diff_us / 2;
        if ((duration_diff_us & 1) && (q & 1)) q += 1;
        m->half_duration_diff = q;
    }
    *out = m;
    return RT_OK;
}

void rt_matcher_destroy(rt_matcher* m) { delete m; }

int rt_matcher_add(rt_matcher* m, const rt_match_signal* sigs, int64_t n) {
    if (!m || (n > 0 && !sigs) || n < 0) return rt_internal_fail(RT_ERR_INVALID, "bad argument");
    for (int64_t i = 0; i < n; ++i) {
        const rt_match_signal& s = sigs[i];
        if (s.device < 0) return rt_internal_fail(RT_ERR_INVALID, "negative device index");
        const int64_t horizon = s.ts_us - m->timeout;       // `now - self.matching_timeout`, now = signal.ts
        const size_t n_open = m->hot.size();
        size_t hit = n_open;
        {
            // the walk: per group 24 bytes and one rarely-taken branch (closed, timed out or member)
            const Hot* const hp = m->hot.data();
            const double f_lo = s.frequency - m->half_bandwidth, f_hi = s.frequency + m->half_bandwidth;
            const int64_t t_lo = s.ts_us - m->time_diff, t_hi = (s.ts_us + s.duration_us) + m->time_diff;
            const int64_t d_lo = s.duration_us - m->half_duration_diff, d_hi = s.duration_us + m->half_duration_diff;
            const bool use_d = m->use_duration;
            for (size_t r = m->head; r < n_open; ++r) {
                const Hot h = hp[r];
                // has_member (__init__.py:337-383) with the signal-side terms hoisted; all comparisons evaluated, branch-free
                unsigned member = (unsigned)!(f_lo > h.frequency) & (unsigned)!(f_hi < h.frequency) &
                                  (unsigned)!(t_lo > h.ts + h.duration) & (unsigned)!(t_hi < h.ts);
                if (use_d) member &= (unsigned)!(d_lo > h.duration) & (unsigned)!(d_hi < h.duration);
                const unsigned special = (unsigned)(h.duration < 0) | (unsigned)(h.ts < horizon);
                if (!(member | special)) continue;
                if (h.duration < 0) continue;                    // closed earlier, not compacted yet
                if (h.ts < horizon) {                            // timed out: consume (match.py:68-71), keep walking
                    emit(*m, m->open[r]);
                    m->open[r].members = std::vector<rt_match_signal>();
                    m->hot[r].duration = -1;
                    ++m->n_dead;
                    continue;
                }
                hit = r;                                         // first fit (match.py:73-76): the walk ends here
                break;
            }
        }
        if (hit != n_open) {
            Group& g = m->open[hit];
            add_member(g, s);
            m->hot[hit] = Hot{g.ts, g.duration, g.frequency};
        }
        // closed groups are mostly the oldest ones: skip the closed prefix, compact (order preserved) only when
        // the closed entries outnumber the open ones -- amortised O(1) per closed group
        while (m->head < n_open && m->hot[m->head].duration < 0) { ++m->head; --m->n_dead; }
        if (m->head + m->n_dead > 64 && 2 * (m->head + m->n_dead) > n_open) {
            size_t w = 0;
            for (size_t r = m->head; r < n_open; ++r) {
                if (m->hot[r].duration < 0) continue;
                if (w != r) { m->open[w] = std::move(m->open[r]); m->hot[w] = m->hot[r]; }
                ++w;
            }
            m->open.resize(w);
            m->hot.resize(w);
            m->head = 0;
            m->n_dead = 0;
        }
        if (hit == n_open) {                                 // match.py:78-81 (n_open: the size before a compaction; only "no hit" matters)
            Group g;
            g.members.push_back(s);
            g.refresh();
            m->hot.push_back(Hot{g.ts, g.duration, g.frequency});
            m->open.push_back(std::move(g));
        }
    }
    return RT_OK;
}

int rt_matcher_pending(const rt_matcher* m, int64_t* n_groups, int64_t* n_members) {
    if (!m) return rt_internal_fail(RT_ERR_INVALID, "null matcher");
    if (n_groups) *n_groups = (int64_t)m->out_sizes.size();
    if (n_members) *n_members = (int64_t)m->out_ids.size();
    return RT_OK;
}

int rt_matcher_drain(rt_matcher* m, int64_t* group_sizes, int64_t* member_ids) {
    if (!m) return rt_internal_fail(RT_ERR_INVALID, "null matcher");
    if ((!m->out_sizes.empty() && !group_sizes) || (!m->out_ids.empty() && !member_ids)) return rt_internal_fail(RT_ERR_INVALID, "null output");
    std::copy(m->out_sizes.begin(), m->out_sizes.end(), group_sizes);
    std::copy(m->out_ids.begin(), m->out_ids.end(), member_ids);
    m->out_sizes.clear();
    m->out_ids.clear();
    return RT_OK;
}

int rt_matcher_open(const rt_matcher* m, int64_t* n_groups, int64_t* n_members) {
    if (!m) return rt_internal_fail(RT_ERR_INVALID, "null matcher");
    int64_t nm = 0, ng = 0;
    for (size_t g = m->head; g < m->open.size(); ++g)
        if (m->hot[g].duration >= 0) { ++ng; nm += (int64_t)m->open[g].members.size(); }
    if (n_groups) *n_groups = ng;
    if (n_members) *n_members = nm;
    return RT_OK;
}

int rt_matcher_read_open(const rt_matcher* m, int64_t* group_sizes, int64_t* member_ids) {
    if (!m) return rt_internal_fail(RT_ERR_INVALID, "null matcher");
    if (m->open.size() > m->head + m->n_dead && (!group_sizes || !member_ids)) return rt_internal_fail(RT_ERR_INVALID, "null output");
    size_t k = 0, w = 0;
    for (size_t g = m->head; g < m->open.size(); ++g) {
        if (m->hot[g].duration < 0) continue;
        group_sizes[w++] = (int64_t)m->open[g].members.size();
        for (const auto& mem : m->open[g].members) member_ids[k++] = mem.id;
    }
    return RT_OK;
}

}  // extern "C"
