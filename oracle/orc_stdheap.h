/*
 * oracle/orc_stdheap.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * The reference orders both its analytic-shot candidates (RS/src/HybridAstar.cpp:216-229) and its A*
 * open list (RS/src/HybridAstar.cpp:589-614) with std::priority_queue and a comparator that looks at
 * the cost only (HybridAstar.h:27-42: lhs.first > rhs.first).  Equal costs are common (every forward
 * Reeds-Shepp segment costs exactly 1.0), so which element pops first is decided by the heap
 * algorithm itself.  This header restates libstdc++'s std::push_heap / std::pop_heap
 * (bits/stl_heap.h: __push_heap, __adjust_heap, __pop_heap) for that comparator so the oracle pops
 * in the same order as a reference built with GCC.
 */
#ifndef ORC_STDHEAP_H
#define ORC_STDHEAP_H

typedef struct {
  double key;
  int payload;
} orc_heap_item;

/* comp(a, b) == (a.key > b.key): "a sinks below b" => min-heap on key */
static void orc__push_heap(orc_heap_item *first, long hole, long top, orc_heap_item value) {
  long parent = (hole - 1) / 2;
  while (hole > top && first[parent].key > value.key) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}

static void orc__adjust_heap(orc_heap_item *first, long hole, long len, orc_heap_item value) {
  const long top = hole;
  long second = hole;
  while (second < (len - 1) / 2) {
    second = 2 * (second + 1);
    if (first[second].key > first[second - 1].key) second--;
    first[hole] = first[second];
    hole = second;
  }
  if ((len & 1) == 0 && second == (len - 2) / 2) {
    second = 2 * (second + 1);
    first[hole] = first[second - 1];
    hole = second - 1;
  }
  orc__push_heap(first, hole, top, value);
}

/* priority_queue::push: c.push_back(x); std::push_heap(c.begin(), c.end(), comp) */
static void orc_stdheap_push(orc_heap_item *h, int *n, orc_heap_item it) {
  h[*n] = it;
  (*n)++;
  orc__push_heap(h, (long)*n - 1, 0, it);
}

/* priority_queue::top + pop: std::pop_heap(c.begin(), c.end(), comp); c.pop_back() */
static orc_heap_item orc_stdheap_pop(orc_heap_item *h, int *n) {
  orc_heap_item top = h[0];
  if (*n > 1) {
    orc_heap_item value = h[*n - 1];
    h[*n - 1] = h[0];
    orc__adjust_heap(h, 0, (long)*n - 1, value);
  }
  (*n)--;
  return top;
}

#endif
