// VecC / SpMatC typedefs (reference: include/edgefem/fem.hpp:11-12, maxwell.hpp:21-22).
// The scalar Helmholtz stub of the reference's fem.hpp is out of scope.
#pragma once
#include "edgefem/linalg.hpp"

namespace edgefem {
using VecC = VectorXcd;
using SpMatC = SparseMatrix<cplx>;
} // namespace edgefem
