"""B200-native belief-propagation message updates behind ITensorNetworksNext.jl's `beliefpropagation` API.

Host side: a Python mirror of the reference interface (beliefpropagation.py, tensornetwork.py, graphs.py)
lowering to the C ABI of csrc/libbpx.so (include/bpx.h).  All arithmetic runs in CUDA kernels for sm_100a.
"""
from . import _lib
from ._lib import BPXError
from . import algorithmsinterface as AI
from . import algorithmsinterface as AIE  # the reference splits the same protocol over two modules
from .algorithmsinterface import MethodError, StopWhenAny, StoppingCriterion
from .beliefpropagation import (
    BeliefPropagationAlgorithm, BeliefPropagationProblem, BeliefPropagationState, BeliefPropagationSweepAlgorithm,
    BeliefPropagationSweepProblem, BeliefPropagationSweepState, DeviceMessageCache, device_iterate,
    ArgumentError, B200MessageUpdate, BeliefPropagationResult, MessageCache, MessageUpdateAlgorithm,
    SimpleMessageUpdate, StopAfterIteration, StopWhenConverged, beliefpropagation, bethe_free_energy,
    default_algorithm, default_beliefpropagation_edges, edge_scalar, edge_scalars, expect, identity_message,
    incoming_messages, iterate_diff, message_environment, message_update, messagecache, ones_message,
    region_scalar, select_algorithm, select_beliefpropagation_stopping_criterion, similar_message_environment,
    vertex_scalar, vertex_scalars,
)
from .apply import (
    ApplyOperatorAlgorithm, BPApplyGate, NoApplyOperatorEnvironmentPreparation, Operator, apply_operator, apply_operators,
    expect_two_site,
)
from .device import BPXContext, fill_randn
from .generators import delta, delta_network, diagonaltensor, ising_network, sqrt_ising_bond
from .graphs import (
    NamedEdge, NamedGraph, forest_cover_edge_sequence, graph_arrays, heavy_hex_127, named_comb_tree,
    named_cycle_graph, named_grid, named_path_graph,
)
from .resident import ResidentState
from .tensornetwork import (
    BraView, Index, ITensor, ITensorNetwork, KetView, NormNetwork, canonical_arrays, normnetwork, random_state,
    randn_itensor, tensornetwork, uniquename,
)
