/*
 * rt_matcher.h -- C ABI of the batched cross-device signal matcher (librtb200.so, host code).
 *
 * First consumer of the detection path's output.  It replaces the inside of
 * radiotracking.match.SignalMatcher.add (radiotracking/match.py:54-82) and the MatchingSignal predicates
 * it calls (radiotracking/__init__.py:285-406: ts = min, duration = max, frequency = statistics.median of the
 * members; has_member; add_member) with an order-preserving native loop over a whole batch of Signals:
 * the reference walks its open groups first-fit in Python for every Signal, which is fine at 5 signals/s and
 * the bottleneck behind a GPU engine that emits 10^5..10^6 signals/s.
 *
 * All times are integer microseconds (Python's datetime/timedelta resolution), so every comparison is exact.
 * The Python binding is pyradiotracking_b200/match.py.  Same conventions as rt_engine.h: plain C types, 0 or a
 * negative RT_ERR_* code, rt_last_error() for the message, one host thread per handle.
 */
#ifndef RT_MATCHER_H
#define RT_MATCHER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One Signal as the matcher sees it (radiotracking/__init__.py:110-170): 48 bytes. */
typedef struct rt_match_signal {
    int64_t ts_us;        /* Signal.ts, microseconds since any fixed epoch */
    int64_t duration_us;  /* Signal.duration */
    double frequency;     /* Signal.frequency [Hz] */
    double avg;           /* Signal.avg [dBW]: decides which of two signals of the same device stays (__init__.py:397-404) */
    int32_t device;       /* index of Signal.device in the caller's device table (any value >= 0) */
    int32_t reserved;
    int64_t id;           /* caller's handle of the Signal object, returned in the emitted groups */
} rt_match_signal;

typedef struct rt_matcher rt_matcher;

/*
 * SignalMatcher.__init__ (match.py:33-50).  timeout_us / time_diff_us are the timedeltas of match.py:41-42 in
 * microseconds; duration_diff_us < 0 means "no duration criterion" (matching_duration_diff_ms None or 0,
 * match.py:44 and __init__.py:376).
 */
int rt_matcher_create(int64_t timeout_us, int64_t time_diff_us, double bandwidth_hz, int64_t duration_diff_us, rt_matcher **out);
void rt_matcher_destroy(rt_matcher *m);

/*
 * SignalMatcher.add (match.py:54-82) for `n` signals in order.  For each one: open groups are visited in creation
 * order; a group whose ts is older than signal.ts - timeout is emitted and closed (match.py:68-71); the first
 * group that has_member() takes the signal (match.py:73-76); otherwise a new group is opened (match.py:78-81).
 * Emitted groups queue up inside the handle until rt_matcher_drain.
 */
int rt_matcher_add(rt_matcher *m, const rt_match_signal *sigs, int64_t n);

/* Emitted-but-not-drained groups and their total member count. */
int rt_matcher_pending(const rt_matcher *m, int64_t *n_groups, int64_t *n_members);
/*
 * Copy out and forget the emitted groups, in emission order (= the order of signal_queue.put, match.py:51):
 * group_sizes[g] members each, member ids concatenated in the order the group's dict holds them
 * (first insertion per device, __init__.py:395-406).
 */
int rt_matcher_drain(rt_matcher *m, int64_t *group_sizes, int64_t *member_ids);

/* The groups still open (SignalMatcher._matched), same encoding, without closing them. */
int rt_matcher_open(const rt_matcher *m, int64_t *n_groups, int64_t *n_members);
int rt_matcher_read_open(const rt_matcher *m, int64_t *group_sizes, int64_t *member_ids);

#ifdef __cplusplus
}
#endif
#endif /* RT_MATCHER_H */
