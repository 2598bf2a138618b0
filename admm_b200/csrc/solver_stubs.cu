// temporary: solvers not yet implemented fail loudly
#include "solvers.h"
namespace b200 {
void solve_wide(const LassoRequest&, b200admm_path*) { throw ArgError("wide (n <= p) solver not implemented yet"); }
void solve_consensus(const LassoRequest&, int, b200admm_path*) { throw ArgError("consensus solver not implemented yet"); }
void solve_lad(const b200admm_data*, bool, const b200admm_opts&, b200admm_dense*) { throw ArgError("lad not implemented yet"); }
void solve_bp(const b200admm_data*, const b200admm_opts&, b200admm_path*) { throw ArgError("bp not implemented yet"); }
}
