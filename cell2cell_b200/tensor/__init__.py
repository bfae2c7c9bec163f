"""Drop-in for the hot-path part of ``cell2cell.tensor`` (cell2cell/tensor/__init__.py:1-8)."""
from .metrics import correlation_index, pairwise_correlation_index
from .tensor import (BaseTensor, InteractionTensor, PreBuiltTensor, build_context_ccc_tensor, generate_tensor_metadata,
                     interactions_to_tensor)
from .factorization import (non_negative_parafac, non_negative_parafac_hals, normalized_error)
from .cp import CPTensor
from .external_scores import dataframes_to_tensor
from .coupled_factorization import coupled_non_negative_parafac
from .coupled_tensor import CoupledInteractionTensor
from .subset import concatenate_interaction_tensors, find_element_indexes, subset_metadata, subset_tensor

__all__ = ["BaseTensor", "InteractionTensor", "PreBuiltTensor", "build_context_ccc_tensor", "generate_tensor_metadata",
           "interactions_to_tensor",
           "correlation_index", "pairwise_correlation_index", "non_negative_parafac", "non_negative_parafac_hals",
           "normalized_error", "CPTensor", "dataframes_to_tensor", "subset_tensor", "subset_metadata", "find_element_indexes",
           "concatenate_interaction_tensors", "coupled_non_negative_parafac", "CoupledInteractionTensor"]
