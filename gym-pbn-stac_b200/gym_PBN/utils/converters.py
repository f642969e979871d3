"""logic_func_data -> PBN_data (front half of the network compiler).

Semantics follow the reference converter (gym_PBN/utils/converters.py:9-40): a node's input mask is the
union of the symbols of all its functions; its table entry for an input assignment is the sum, in function
order, of the probabilities of the functions that evaluate to 1; a node without inputs is a control node.
"""
import itertools

import numpy as np

from .logic.eval import LogicExpressionEvaluator, compile_expression


def logic_funcs_to_PBN_data(nodes, node_functions):
    n = len(nodes)
    position = {}
    for j, name in enumerate(nodes):
        position.setdefault(name, j)  # list.index semantics: first occurrence
    pbn_data = []
    for i, name in enumerate(nodes):
        funcs = [(compile_expression(expr), prob) for expr, prob in node_functions[i]]
        mask = np.zeros(n, dtype=bool)
        for expr, _ in node_functions[i]:
            for sym in LogicExpressionEvaluator.get_symbols(expr):
                if sym not in position:
                    raise ValueError(f"'{sym}' is not in list")
                mask[position[sym]] = True
        inputs = [nodes[j] for j in np.nonzero(mask)[0]]
        table = np.zeros([2] * len(inputs))
        for assignment in itertools.product([0, 1], repeat=len(inputs)):
            env = dict(zip(inputs, assignment))
            for fn, prob in funcs:
                if int(fn(env)) == 1:
                    table[assignment] += prob
        pbn_data.append((mask, table, name, len(inputs) == 0))
    return pbn_data
