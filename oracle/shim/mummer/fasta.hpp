// TEST INFRASTRUCTURE ONLY (oracle build shim): empty, see sparseSA.hpp.
