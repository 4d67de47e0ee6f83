"""Stand-in for the part of `hippopt.OptiSolver` the shim inherits (fields and helpers of base/opti_solver.py:49-229)."""


class OptiFailure(Exception):
    def __init__(self, message, callback_used):
        info = " and the callback did not manage to save an intermediate solution" if callback_used else ""
        super().__init__(f"Opti failed to solve the problem{info}. Message: {str(message)}")


class OptiSolver:
    def __init__(self, opti, options_plugin=None, options_solver=None, inner_solver="ipopt"):
        self._solver = opti
        self._inner_solver = inner_solver
        self._options_plugin, self._options_solver = dict(options_plugin or {}), dict(options_solver or {})
        self._cost = None
        self._cost_expressions, self._constraint_expressions = {}, {}
        self._free_parameters = []
        self._variables_map, self._parameters_map = {}, {}
        self._objects = "objects"
        self._callback_criterion = None
        self._output_cost = self._output_solution = self._cost_values = self._constraint_values = None
        self.stock_solve_called = False

    def _generate_solution_output(self, variables, input_solution):
        out = {self._variables_map[v]: input_solution[v] for v in self._variables_map}
        out.update({self._parameters_map[v]: input_solution[v] for v in self._parameters_map})
        return out

    def solve(self):
        self.stock_solve_called = True
