/* oracle/rand_shim.c -- TEST INFRASTRUCTURE ONLY.
 * Linked into oracle/_ref/ez_tree*.so (see build_ref.py): the EfficientZero reference tree breaks PUCT ties with
 * rand() reseeded from the wall clock (ctree_efficientzero/lib/cnode.cpp:691, common_lib/utils.cpp:12-26) and has no
 * deterministic switch.  These hidden-visibility definitions bind the module's own rand()/srand() calls at link time
 * (nothing is exported, libc's rand is untouched for everybody else) so that the unmodified sources always take
 * element 0 of the tie list = the first child attaining the exact maximum. */
__attribute__((visibility("hidden"))) int rand(void) { return 0; }
__attribute__((visibility("hidden"))) void srand(unsigned seed) { (void)seed; }
