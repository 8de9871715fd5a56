// Order-exact emulation of torch.topk(x, k, dim=0, largest=True, sorted=True) on CPU for one column.
//
// PMVO.Find_max_conf_from_visible_view (PMVO.py:339-343) takes torch.topk(Conf', 20) over the views.  Conf' is
// quantised (k/255, and exactly 0 for invisible views) so equal values are the common case, and which of two
// equal views lands on an even rank decides the base views PMVO.forward uses (PMVO.py:50-51).  The order among
// equal values is not specified by torch; on CPU it is whatever ATen's kernel produces:
//     queue[j] = (x[j], j);  if (k*64 <= n) partial_sort  else  { nth_element(k-1); sort(first k-1) }
// with comp(a,b) = (isnan(a) && !isnan(b)) || a > b  (aten/src/ATen/native/cpu/TopKImpl.h), i.e. libstdc++'s
// introselect / introsort.  Those two algorithms are re-implemented here from their published description
// (median-of-3 to first, unguarded Hoare partition, insertion sort below the thresholds 3 / 16, heap-select
// fallback at depth 0) so the GPU picks the same base views as the CPU reference.  tests/test_topk_host.py
// checks it against torch.topk on tie-heavy inputs through mh_debug_topk_host.
#pragma once
#include "mh_common.cuh"

struct MhKV { float v; int i; };

MH_HD bool mh_tk_less(const MhKV& a, const MhKV& b) {       // "a orders before b" (descending, NaN first)
    bool an = a.v != a.v, bn = b.v != b.v;
    return (an && !bn) || (a.v > b.v);
}
MH_HD void mh_tk_swap(MhKV* q, int a, int b) { MhKV t = q[a]; q[a] = q[b]; q[b] = t; }

MH_HD void mh_tk_median_to_first(MhKV* q, int result, int a, int b, int c) {
    if (mh_tk_less(q[a], q[b])) {
        if (mh_tk_less(q[b], q[c])) mh_tk_swap(q, result, b);
        else if (mh_tk_less(q[a], q[c])) mh_tk_swap(q, result, c);
        else mh_tk_swap(q, result, a);
    } else if (mh_tk_less(q[a], q[c])) mh_tk_swap(q, result, a);
    else if (mh_tk_less(q[b], q[c])) mh_tk_swap(q, result, c);
    else mh_tk_swap(q, result, b);
}
MH_HD int mh_tk_partition(MhKV* q, int first, int last, int pivot) {
    for (;;) {
        while (mh_tk_less(q[first], q[pivot])) ++first;
        --last;
        while (mh_tk_less(q[pivot], q[last])) --last;
        if (!(first < last)) return first;
        mh_tk_swap(q, first, last);
        ++first;
    }
}
MH_HD int mh_tk_partition_pivot(MhKV* q, int first, int last) {
    int mid = first + (last - first) / 2;
    mh_tk_median_to_first(q, first, first + 1, mid, last - 1);
    return mh_tk_partition(q, first + 1, last, first);
}
MH_HD void mh_tk_linear_insert(MhKV* q, int last) {
    MhKV val = q[last];
    int next = last - 1;
    while (mh_tk_less(val, q[next])) { q[last] = q[next]; last = next; --next; }
    q[last] = val;
}
MH_HD void mh_tk_insertion_sort(MhKV* q, int first, int last) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (mh_tk_less(q[i], q[first])) {
            MhKV val = q[i];
            for (int j = i; j > first; --j) q[j] = q[j - 1];
            q[first] = val;
        } else mh_tk_linear_insert(q, i);
    }
}
// ---- heap helpers (depth-limit fallback) ----
MH_HD void mh_tk_push_heap(MhKV* q, int first, int hole, int top, MhKV val) {
    int parent = (hole - 1) / 2;
    while (hole > top && mh_tk_less(q[first + parent], val)) {
        q[first + hole] = q[first + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    q[first + hole] = val;
}
MH_HD void mh_tk_adjust_heap(MhKV* q, int first, int hole, int len, MhKV val) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (mh_tk_less(q[first + child], q[first + child - 1])) --child;
        q[first + hole] = q[first + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        q[first + hole] = q[first + child - 1];
        hole = child - 1;
    }
    mh_tk_push_heap(q, first, hole, top, val);
}
MH_HD void mh_tk_make_heap(MhKV* q, int first, int last) {
    int len = last - first;
    if (len < 2) return;
    int parent = (len - 2) / 2;
    for (;;) {
        MhKV val = q[first + parent];
        mh_tk_adjust_heap(q, first, parent, len, val);
        if (parent == 0) return;
        --parent;
    }
}
MH_HD void mh_tk_pop_heap(MhKV* q, int first, int last, int result) {
    MhKV val = q[result];
    q[result] = q[first];
    mh_tk_adjust_heap(q, first, 0, last - first, val);
}
MH_HD void mh_tk_heap_select(MhKV* q, int first, int middle, int last) {
    mh_tk_make_heap(q, first, middle);
    for (int i = middle; i < last; ++i)
        if (mh_tk_less(q[i], q[first])) mh_tk_pop_heap(q, first, middle, i);
}
MH_HD void mh_tk_sort_heap(MhKV* q, int first, int last) {
    while (last - first > 1) { --last; mh_tk_pop_heap(q, first, last, last); }
}
MH_HD int mh_tk_lg(int n) { int k = 0; while (n > 1) { n >>= 1; ++k; } return k; }

MH_HD void mh_tk_nth_element(MhKV* q, int first, int nth, int last) {
    if (first == last || nth == last) return;
    int depth = mh_tk_lg(last - first) * 2;
    while (last - first > 3) {
        if (depth == 0) {
            mh_tk_heap_select(q, first, nth + 1, last);
            mh_tk_swap(q, first, nth);
            return;
        }
        --depth;
        int cut = mh_tk_partition_pivot(q, first, last);
        if (cut <= nth) first = cut; else last = cut;
    }
    mh_tk_insertion_sort(q, first, last);
}
// std::sort: introsort loop is recursive on the right part; ranges here are <= k-1 elements so an explicit
// stack of a few entries is enough (depth limit 2*lg(n)).
MH_HD void mh_tk_sort(MhKV* q, int first, int last) {
    if (first == last) return;
    int st_first[40], st_last[40], st_depth[40], sp = 0;
    st_first[0] = first; st_last[0] = last; st_depth[0] = mh_tk_lg(last - first) * 2; sp = 1;
    while (sp > 0) {
        --sp;
        int f = st_first[sp], l = st_last[sp], d = st_depth[sp];
        while (l - f > 16) {
            if (d == 0) {                      // partial_sort(f, l, l) = heap sort
                mh_tk_heap_select(q, f, l, l);
                mh_tk_sort_heap(q, f, l);
                break;
            }
            --d;
            int cut = mh_tk_partition_pivot(q, f, l);
            // libstdc++ recurses on [cut,l) first, then loops on [f,cut): order of the two is irrelevant
            // because the ranges are disjoint.
            if (sp < 40) { st_first[sp] = cut; st_last[sp] = l; st_depth[sp] = d; ++sp; }
            l = cut;
        }
    }
    // final insertion sort
    if (last - first > 16) {
        mh_tk_insertion_sort(q, first, first + 16);
        for (int i = first + 16; i != last; ++i) mh_tk_linear_insert(q, i);
    } else mh_tk_insertion_sort(q, first, last);
}
MH_HD void mh_tk_partial_sort(MhKV* q, int first, int middle, int last) {
    mh_tk_heap_select(q, first, middle, last);
    mh_tk_sort_heap(q, first, middle);
}

// q[0..n) initialised to (x[j], j); afterwards q[0..k) is torch.topk's (values, indices).
MH_HD void mh_topk_torch_cpu(MhKV* q, int n, int k) {
    if (k <= 0) return;
    if (k > n) k = n;
    if ((long long)k * 64 <= n) {
        mh_tk_partial_sort(q, 0, k, n);
    } else {
        mh_tk_nth_element(q, 0, k - 1, n);
        mh_tk_sort(q, 0, k - 1);
    }
}
