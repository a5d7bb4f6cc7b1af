// Host-side batch scheduler: turns the per-trajectory jump events of one batch into the sequence of kernel launches
// (tile launches over Work items, decide launches over DecideItems).  Pure C++; shared by the CUDA engine and by the
// CPU emulation used in tests.
//
// All trajectories of a batch walk the program's segments in lock step.  A trajectory with an event at instruction e
// inside segment s runs s in pieces: [.., e] with the bare gate (and, if the branch depends on the state, the
// reduced-density-matrix reduction), then -- after the decide step -- the chosen Kraus operator followed by
// [e+1, ..).  The pieces after the first are issued as extra, small launches ("rounds") over the affected slots only.
#pragma once
#include <vector>

#include "ncb_compile.h"

namespace ncb {

struct LaunchStep {
    int kind;          // 0: tile launch over works[begin, begin+count); 1: decide launch over decides[begin, begin+count)
    int begin, count;
    int seg;           // tile launches: segment (all works of a launch share it)
};
struct BatchPlan {
    std::vector<Work> works;
    std::vector<DecideItem> decides;
    std::vector<LaunchStep> steps;
    long long sweeps = 0;          // slot-sweeps issued (works)
};

// events[slot]: that trajectory's events in instruction order.  init: the batch starts from |0..0> (first segment
// does not read the state).  want_norm: emit WF_NORM on every slot's final work.
inline void schedule_batch(const Program& P, const std::vector<std::vector<Event>>& events, bool init, bool want_norm,
                           BatchPlan& out) {
    out.works.clear(); out.decides.clear(); out.steps.clear(); out.sweeps = 0;
    const int B = (int)events.size(), S = (int)P.segs.size();
    std::vector<size_t> cur(B, 0);
    std::vector<int> live, next;
    for (int s = 0; s < S; ++s) {
        const Segment& sg = P.segs[s];
        const bool first = init && s == 0, last = s == S - 1;
        // events of this segment per slot: [cur[b], end[b])
        std::vector<size_t> end(B);
        for (int b = 0; b < B; ++b) {
            size_t e = cur[b];
            while (e < events[b].size() && events[b][e].instr < sg.instr_hi) ++e;
            end[b] = e;
        }
        // main launch
        LaunchStep st{0, (int)out.works.size(), B, s};
        live.clear();
        for (int b = 0; b < B; ++b) {
            Work w{b, s, sg.instr_lo, sg.instr_hi, first ? WF_INIT : 0, -1, -1, -1};
            if (cur[b] < end[b]) {
                const Event& e = events[b][cur[b]];
                w.instr_hi = e.instr + 1;
                w.flags |= WF_BARE_LAST;
                if (e.kmin != e.kmax) { w.flags |= WF_REDUCE; w.red_instr = e.instr; }
                live.push_back(b);
            } else if (last && want_norm) {
                w.flags |= WF_NORM;
            }
            out.works.push_back(w);
        }
        out.steps.push_back(st);
        out.sweeps += B;
        // rounds
        while (!live.empty()) {
            LaunchStep ds{1, (int)out.decides.size(), 0, s};
            for (int b : live) {
                const Event& e = events[b][cur[b]];
                if (e.kmin != e.kmax) { out.decides.push_back(DecideItem{b, e.instr, e.u}); ++ds.count; }
            }
            if (ds.count) out.steps.push_back(ds);
            LaunchStep ts{0, (int)out.works.size(), (int)live.size(), s};
            next.clear();
            for (int b : live) {
                const Event& e = events[b][cur[b]];
                ++cur[b];
                Work w{b, s, e.instr + 1, sg.instr_hi, 0, e.instr, e.kmin, -1};
                w.flags |= (e.kmin == e.kmax) ? WF_PRO_STATIC : WF_PRO_DYNAMIC;
                if (cur[b] < end[b]) {
                    const Event& f = events[b][cur[b]];
                    w.instr_hi = f.instr + 1;
                    w.flags |= WF_BARE_LAST;
                    if (f.kmin != f.kmax) { w.flags |= WF_REDUCE; w.red_instr = f.instr; }
                    next.push_back(b);
                } else if (last && want_norm) {
                    w.flags |= WF_NORM;
                }
                out.works.push_back(w);
            }
            out.steps.push_back(ts);
            out.sweeps += (long long)live.size();
            live.swap(next);
        }
    }
}

}  // namespace ncb
