#pragma once
#include <stdint.h>
#include <vector>

namespace nepb {

struct UnionCSR {
    int64_t n = 0, nnz = 0;
    int p = 0, vw = 0;
    std::vector<int64_t> colptr;      // CSC union, 0-based
    std::vector<int32_t> rowval;      // CSC union
    std::vector<int32_t> rowptr;      // CSR union
    std::vector<int32_t> colind;      // CSR union
    std::vector<int32_t> csr_of_csc;  // CSC position -> CSR position
    std::vector<double> vals;         // [nnz][vw] in CSR order
};

int build_union_csr(int64_t n, int p, const int64_t* const* colptr, const int64_t* const* rowval,
                    const void* const* nzval, int is_complex, int base, UnionCSR& u);

}  // namespace nepb
