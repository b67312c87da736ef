// TEST INFRASTRUCTURE ONLY (oracle build shim): select support is included but never used by the alignment path.
