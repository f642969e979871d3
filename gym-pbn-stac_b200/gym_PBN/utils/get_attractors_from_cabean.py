"""Attractor lists in the CABEAN interchange format (reference: gym_PBN/utils/get_attractors_from_cabean.py).

The reference shells out to the external CABEAN binary (utils/get_cabean_model.py:95, tool and model template not
shipped) and parses its report into one list of cubes per attractor, a cube being a tuple over {0, 1, '*'}.  Here
`parse_state` / `parse_attractors` read the same report text (for lists computed elsewhere), and `get_attractors(env)`
produces the list without the tool: the exhaustive terminal-SCC search on the GPU for networks up to 32 genes
(gym_PBN.b200.attractors), the reference's sampling recipe beyond that.
"""
import pickle


def parse_state(spec):
    """One report line's state: symbols at the even positions, '-' = don't care (:9-11)."""
    return tuple("*" if v == "-" else int(v) for v in spec[0:len(spec):2])


def parse_attractors(cabean_out):
    """{attractor number (0-based): [cube, ...]} from a CABEAN report (:14-36)."""
    attractors = {}
    num = None
    for line in cabean_out.split("\n"):
        if line.startswith("=") and "=== find attractor #" in line:
            num = int(line.split()[3][1:]) - 1
        elif num is not None:
            if line.startswith(":"):
                continue
            if not line:
                num = None
                continue
            attractors.setdefault(num, []).append(parse_state(line.split()[0]))
    return attractors


def get_attractors(env, cache=None):
    """List of attractors (each a list of cubes) of env's network, optionally pickled to `cache` as the reference does
    with data/attractors_{env.name}.pkl (:39-54)."""
    from gym_PBN.b200 import attractors as att_tools

    core = getattr(env, "unwrapped", env)
    care = getattr(core, "target_node_indices", None) or None
    attractors, _source = att_tools.default_attractors(core.network, care, seed=0, exact_max_nodes=32)
    if cache is not None:
        with open(cache, "wb+") as f:
            pickle.dump(attractors, f)
    return attractors


# the report the reference keeps "for testing" (:57-82); tests/test_host_logic.py checks the parser against it
sample_cabean_out = r"""***************************************************************************
                       CABEAN 2.0.0 
 Please check http://satoss.uni.lu/software/CABEAN/ for the latest release.
 Please send any feedback to <cui.su@uni.lu>
***************************************************************************

Command line: cabean model_from_jinja.ispl
======================== find attractor #1 : 4 states ========================
: 6 nodes 1 leaves 4 minterms
1-0-1-0-----1-  1

======================== find attractor #2 : 1 states ========================
: 8 nodes 1 leaves 1 minterms
1-0-1-1-1-1-0-  1

======================== find attractor #3 : 1 states ========================
: 8 nodes 1 leaves 1 minterms
1-0-1-1-1-1-1-  1

======================== find attractor #4 : 1 states ========================
: 8 nodes 1 leaves 1 minterms
1-1-1-1-1-1-0-  1

number of attractors = 4
time for attractor detection=0
"""
