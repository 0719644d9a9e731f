// oracle/stl_probe.cpp -- TEST INFRASTRUCTURE ONLY.
// Calls the real libstdc++ std::partial_sort with the comparator the reference uses
// (src/cca.cpp:179-185, :225-228) so the hand-restated heap-select (slic_oracle.c
// orc_heap_select, and the device kernel) can be differential-tested against it.
#include <algorithm>
#include <vector>
extern "C" void stl_partial_sort_by_area(int* comps, long n, long middle, const int* area) {
    struct cmp {
        const int* a;
        bool operator()(int l, int r) const { return a[l] > a[r]; }
    } c{area};
    std::partial_sort(comps, comps + middle, comps + n, c);
}
