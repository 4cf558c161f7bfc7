
#include "action_axpy.hh"
#include "action_axpby.hh"

#include "action_matrix_vector_product.hh"
#include "action_atv_product.hh"

#include "action_matrix_matrix_product.hh"
// #include "action_ata_product.hh"
#include "action_aat_product.hh"

#include "action_trisolve.hh"
#include "action_trmm.hh"
#include "action_symv.hh"
// #include "action_symm.hh"
#include "action_syr2.hh"
#include "action_ger.hh"
#include "action_rot.hh"

// #include "action_lu_solve.hh"

