# oracle build of the reference trimem.core (test infrastructure only)
