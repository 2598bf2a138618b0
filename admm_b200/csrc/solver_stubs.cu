// (all solvers are implemented; this file intentionally left without definitions)
