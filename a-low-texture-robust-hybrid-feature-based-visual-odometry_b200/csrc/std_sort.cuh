// libstdc++'s std::sort, restated so that device code reproduces the ORDER the reference gets from it.
//
// The reference sorts KeyLines with std::sort and a comparator on the response only (include/auxiliar.h:47-52, used at
// src/LineExtractor.cpp:353 and src/Frame.cc:1087).  std::sort is not stable: where two responses are equal, the order of
// the two lines (hence their class_id, which lines survive the nLSDFeature cut, and the row order of the LBD matrix) is
// whatever libstdc++'s introsort leaves.  That order depends only on the sequence of comparator results, so sorting an
// index array with the same algorithm gives the same permutation as sorting the KeyLine structs.
//
// Algorithm (libstdc++ bits/stl_algo.h, bits/stl_heap.h; unchanged since GCC 4.x): introsort loop (median of first+1 / mid /
// last-1 moved to first, unguarded Hoare partition, recursion on the right part, depth limit 2*floor(log2 n), heap sort
// below it) down to ranges of 16, then one insertion sort pass (guarded for the first 16 elements, unguarded after).
//
// `less(a, b)` is the comparator on ELEMENT VALUES (here: indices into a key array).  Host + device; the host build is
// what tests/cpp/std_sort_main.cpp checks against the real std::sort.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define HVO_HD __host__ __device__ __forceinline__
#else
#define HVO_HD inline
#endif

namespace hvo {
namespace stdsort {

template <class T> HVO_HD void iter_swap(T* a, T* b) { const T t = *a; *a = *b; *b = t; }

template <class T, class Less>
HVO_HD void unguarded_linear_insert(T* last, Less less) {
    const T val = *last;
    T* next = last - 1;
    while (less(val, *next)) { *last = *next; last = next; --next; }
    *last = val;
}

template <class T, class Less>
HVO_HD void insertion_sort(T* first, T* last, Less less) {
    if (first == last) return;
    for (T* i = first + 1; i != last; ++i) {
        if (less(*i, *first)) {
            const T val = *i;
            for (T* p = i; p != first; --p) *p = *(p - 1);   // move_backward(first, i, i + 1)
            *first = val;
        } else {
            unguarded_linear_insert(i, less);
        }
    }
}

template <class T, class Less>
HVO_HD void push_heap(T* first, int hole, int top, T value, Less less) {
    int parent = (hole - 1) / 2;
    while (hole > top && less(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

template <class T, class Less>
HVO_HD void adjust_heap(T* first, int hole, int len, T value, Less less) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (less(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    push_heap(first, hole, top, value, less);
}

template <class T, class Less>
HVO_HD void heap_sort(T* first, T* last, Less less) {   // std::__partial_sort(first, last, last): make_heap + sort_heap
    const int len = (int)(last - first);
    if (len >= 2) {
        int parent = (len - 2) / 2;
        while (true) {
            const T value = first[parent];
            adjust_heap(first, parent, len, value, less);
            if (parent == 0) break;
            parent--;
        }
    }
    while (last - first > 1) {
        --last;
        const T value = *last;
        *last = *first;
        adjust_heap(first, 0, (int)(last - first), value, less);
    }
}

template <class T, class Less>
HVO_HD void move_median_to_first(T* result, T* a, T* b, T* c, Less less) {
    if (less(*a, *b)) {
        if (less(*b, *c)) iter_swap(result, b);
        else if (less(*a, *c)) iter_swap(result, c);
        else iter_swap(result, a);
    } else if (less(*a, *c)) iter_swap(result, a);
    else if (less(*b, *c)) iter_swap(result, c);
    else iter_swap(result, b);
}

template <class T, class Less>
HVO_HD T* unguarded_partition(T* first, T* last, T* pivot, Less less) {
    while (true) {
        while (less(*first, *pivot)) ++first;
        --last;
        while (less(*pivot, *last)) --last;
        if (!(first < last)) return first;
        iter_swap(first, last);
        ++first;
    }
}

// std::sort(first, first + n, less).  The recursion of __introsort_loop (on the right part, the loop continues on the left
// part) is kept on an explicit stack: at most depth_limit <= 2 * 31 pending right parts.
template <class T, class Less>
HVO_HD void sort(T* first, int n, Less less) {
    if (n <= 0) return;
    T* last = first + n;
    int lg = 0;
    for (unsigned v = (unsigned)n; v > 1; v >>= 1) ++lg;   // std::__lg
    struct Pending { T* first; T* last; int depth; };
    Pending stack[64];
    int sp = 0;
    T* f = first;
    T* l = last;
    int depth = 2 * lg;
    while (true) {
        bool done = false;
        while (l - f > 16) {
            if (depth == 0) { heap_sort(f, l, less); done = true; break; }
            --depth;
            T* mid = f + (l - f) / 2;
            move_median_to_first(f, f + 1, mid, l - 1, less);
            T* cut = unguarded_partition(f + 1, l, f, less);
            stack[sp].first = cut; stack[sp].last = l; stack[sp].depth = depth; ++sp;   // std::__introsort_loop(cut, last, depth)
            l = cut;
        }
        (void)done;
        if (sp == 0) break;
        --sp;
        f = stack[sp].first; l = stack[sp].last; depth = stack[sp].depth;
    }
    // __final_insertion_sort
    if (n > 16) {
        insertion_sort(first, first + 16, less);
        for (T* i = first + 16; i != last; ++i) unguarded_linear_insert(i, less);
    } else {
        insertion_sort(first, last, less);
    }
}

}  // namespace stdsort
}  // namespace hvo
