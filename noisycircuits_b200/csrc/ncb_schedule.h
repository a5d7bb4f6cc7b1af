// Host-side batch scheduler: turns the per-trajectory jump events of one batch into the sequence of kernel launches
// (tile launches over Work items, decide launches over DecideItems).  Pure C++; shared by the CUDA engine and by the
// CPU emulation used in tests.
//
// All trajectories of a batch walk the program's segments in lock step.  A trajectory with an event at instruction e
// inside segment s runs s in pieces: [.., e] with the bare gate (and, if the branch depends on the state, the
// reduced-density-matrix reduction), then -- after the decide step -- the chosen Kraus operator followed by
// [e+1, ..).  The pieces after the first are issued as extra, small launches ("rounds") over the affected slots only.
#pragma once
#include <algorithm>
#include <vector>

#include "ncb_compile.h"

namespace ncb {

struct LaunchStep {
    int kind;          // 0: tile launch over works[begin, begin+count); 1: decide launch over decides[begin, begin+count);
                       // 2: generic-unitary launch over works[begin, begin+count) (a segment that is one 5..8-bit unitary)
    int begin, count;
    int seg;           // tile launches: segment (all works of a launch share it)
    int direct;        // tile launches: 1 = every work runs the whole segment without an event (direct kernel)
};
struct BatchPlan {
    std::vector<Work> works;
    std::vector<DecideItem> decides;
    std::vector<LaunchStep> steps;
    long long sweeps = 0;          // slot-sweeps issued (works)
    // state slots the plan uses: [0, B) one per trajectory, then (shared prefix only) two trunk slots
    int nslots = 0;
    int trunk_slot = -1;           // slot holding the no-jump trajectory's final state (-1: none)
    std::vector<double> weight;    // [nslots] trajectories whose final state is this slot's (0: slot unused)
};

// events[slot]: that trajectory's events in instruction order.  init: the batch starts from |0..0> (first segment
// does not read the state).  want_norm: emit WF_NORM on every slot's final work.
//
// share_prefix: every trajectory is the SAME deterministic evolution until its first jump event, so that prefix is
// computed once per batch.  A "trunk" walks all segments (ping-ponging between the two extra slots B and B+1, so a
// launch never reads a slot it writes); a trajectory joins at the segment of its first event with its tiles READ from
// the trunk's state at the start of that segment (Work::src) and written to its own slot; trajectories without any
// event are the trunk (weight = their number).  The numbers produced per trajectory are bit-identical to the unshared
// schedule: the same kernels run on the same data, fewer times.
// direct_norm: works of the last segment that only carry WF_NORM count as "whole segment, no event" too.
inline void schedule_batch(const Program& P, const std::vector<std::vector<Event>>& events, bool init, bool want_norm,
                           BatchPlan& out, bool share_prefix = false, bool inline_certain = true, bool direct_norm = false) {
    out.works.clear(); out.decides.clear(); out.steps.clear(); out.sweeps = 0;
    const int B = (int)events.size(), S = (int)P.segs.size();
    share_prefix = share_prefix && init && B > 1 && S > 0;
    out.nslots = share_prefix ? B + 2 : B;
    out.trunk_slot = -1;
    out.weight.assign(out.nslots, share_prefix ? 0.0 : 1.0);
    // segment of each trajectory's first event (S: none)
    std::vector<int> first_seg(B, share_prefix ? S : 0);
    int trunk_until = -1;                     // the trunk has to produce the state at the start of segment trunk_until
    if (share_prefix) {
        int never = 0;
        for (int b = 0; b < B; ++b) {
            if (events[b].empty()) { ++never; continue; }
            int s = 0;
            while (s < S && P.segs[s].instr_hi <= events[b][0].instr) ++s;
            first_seg[b] = s;
            trunk_until = std::max(trunk_until, s);
            out.weight[b] = 1.0;
        }
        if (never) {
            trunk_until = S;
            out.trunk_slot = B + (S & 1);
            out.weight[out.trunk_slot] = (double)never;
        }
    }
    std::vector<size_t> cur(B, 0);
    std::vector<int> live, next;
    for (int s = 0; s < S; ++s) {
        const Segment& sg = P.segs[s];
        const bool first = init && s == 0, last = s == S - 1;
        const int tin = B + (s & 1), tout = B + ((s + 1) & 1);
        // events of this segment per slot: [cur[b], end[b])
        std::vector<size_t> end(B);
        for (int b = 0; b < B; ++b) {
            size_t e = cur[b];
            while (e < events[b].size() && events[b][e].instr < sg.instr_hi) ++e;
            end[b] = e;
        }
        // End of a piece that starts at w.instr_lo: it runs to the trajectory's next event in this segment (returns true:
        // a round follows) or to the end of the segment.  A jump whose branch is known from the uniform alone
        // (kmin == kmax) needs no reduction, so ONE such jump per piece is applied in-line (WF_MID) instead of costing the
        // trajectory another sweep.
        auto finish_piece = [&](Work& w, int b) -> bool {
            if (inline_certain && cur[b] < end[b] && events[b][cur[b]].kmin == events[b][cur[b]].kmax && events[b][cur[b]].kmin < 0x8000) {
                const Event& m = events[b][cur[b]];
                w.flags |= WF_MID | (m.kmin << 16);
                if (end[b] - cur[b] == 1 && w.instr_lo == sg.instr_lo) w.flags |= WF_PULL;      // the only event of the segment
                w.mid_instr = m.instr;
                ++cur[b];
            }
            if (cur[b] < end[b]) {
                const Event& e = events[b][cur[b]];
                w.instr_hi = e.instr + 1;
                w.flags |= WF_BARE_LAST;
                if (e.kmin != e.kmax) w.flags |= WF_REDUCE;
                return true;
            }
            if (last && want_norm) w.flags |= WF_NORM;
            return false;
        };
        // main launch
        LaunchStep st{sg.pad[0] ? 2 : 0, (int)out.works.size(), 0, s, 0};
        live.clear();
        if (share_prefix && s < trunk_until) {
            Work w{tout, tin, sg.instr_lo, sg.instr_hi, first ? WF_INIT : 0, -1, -1, -1};
            if (last && want_norm) w.flags |= WF_NORM;
            out.works.push_back(w);
        }
        for (int b = 0; b < B; ++b) {
            if (first_seg[b] > s) continue;                       // still on the trunk
            Work w{b, first_seg[b] == s && share_prefix ? tin : b, sg.instr_lo, sg.instr_hi, first ? WF_INIT : 0, -1, -1, -1};
            if (finish_piece(w, b)) live.push_back(b);
            out.works.push_back(w);
        }
        st.count = (int)out.works.size() - st.begin;
        out.sweeps += st.count;
        // works that run the whole segment without an event go to the direct kernel where the segment allows it
        const bool seg_direct = (size_t)s < P.segconst.size() && P.segconst[s].valid && P.segconst[s].direct;
        if (seg_direct && st.count) {
            auto first = out.works.begin() + st.begin;
            // (direct_norm: the whole-segment kernel of the LAST segment also accumulates the squared norm -- generated kernels)
            auto mid = std::stable_partition(first, out.works.end(), [&](const Work& w) {
                return (w.flags == 0 || (direct_norm && last && w.flags == WF_NORM)) && w.instr_lo <= sg.instr_lo && w.instr_hi >= sg.instr_hi; });
            const int nd = (int)(mid - first);
            if (nd) out.steps.push_back(LaunchStep{0, st.begin, nd, s, 1});
            if (st.count - nd) out.steps.push_back(LaunchStep{0, st.begin + nd, st.count - nd, s, 0});
        } else if (st.count) {
            out.steps.push_back(st);
        }
        // rounds
        while (!live.empty()) {
            LaunchStep ds{1, (int)out.decides.size(), 0, s, 0};
            for (int b : live) {
                const Event& e = events[b][cur[b]];
                if (e.kmin != e.kmax) { out.decides.push_back(DecideItem{b, e.instr, e.u}); ++ds.count; }
            }
            if (ds.count) out.steps.push_back(ds);
            LaunchStep ts{0, (int)out.works.size(), (int)live.size(), s, 0};
            next.clear();
            for (int b : live) {
                const Event& e = events[b][cur[b]];
                ++cur[b];
                Work w{b, b, e.instr + 1, sg.instr_hi, 0, e.instr, e.kmin, -1};
                w.flags |= (e.kmin == e.kmax) ? WF_PRO_STATIC : WF_PRO_DYNAMIC;
                if (finish_piece(w, b)) next.push_back(b);
                out.works.push_back(w);
            }
            out.steps.push_back(ts);
            out.sweeps += (long long)live.size();
            live.swap(next);
        }
    }
}

}  // namespace ncb
