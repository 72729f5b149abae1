"""B200-native (sm_100a) implementation of the metric-learning hot path of johndpope/multimodal_similarity.

Function-level drop-ins for the reference's leaf functions (same names, argument meaning and return tuples):

    distance   all_diffs / cdist / all_diffs_tf / cdist_tf / pairwise_distance     src/utils.py:302-360
               project_normalize (xw_plus_b + l2_normalize embedding head)         src/networks.py:376-380
    losses     batch_hard / lifted_loss (fused forward + backward)                 src/networks.py:797-870
               triplet_semihard_loss / lifted_struct_loss (tf.contrib equivalents) src/base_CUB.py:163-171
    retrieval  retrieve / retrieve_host / retrieve_one / evaluate / evaluate_simple /
               recall_at_K / precision_at_recall / late_fusion                     src/utils.py:55-266
    mining     select_triplets_facenet (semi-hard negatives counted and picked on device)  src/utils.py:430-496
    sharded    ShardedGallery (gallery rows split over ranks, NCCL merge)          (new)

Everything runs through the C-ABI library libmmsim.so (include/mmsim.h); there is no CPU fallback.
"""
from ._lib import MmsimError, load  # noqa: F401
from .distance import all_diffs, all_diffs_tf, cdist, cdist_tf, pairwise_distance, project_normalize  # noqa: F401
from .losses import batch_hard, lifted_loss, lifted_struct_loss, triplet_semihard_loss  # noqa: F401
from .retrieval import (  # noqa: F401
    average_precision, evaluate, evaluate_simple, full_ranking, late_fusion, precision_at_recall, recall_at_K, retrieve,
    retrieve_host, retrieve_one,
)
from .mining import select_triplets_facenet, select_triplets_facenet_cub, semihard_counts  # noqa: F401
from .sharded import ShardedGallery  # noqa: F401

__version__ = "0.1.0"
