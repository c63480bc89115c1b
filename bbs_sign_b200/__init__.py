"""bbs_sign_b200 -- B200-native (sm_100a) batch engine for the BBS verification hot path of
hashcloak/bbs_sign.  See DESIGN.md.  The CUDA library is built in-tree by `__graft_entry__.build()`;
nothing here falls back to a CPU implementation."""
from .api import (BLS12_381, BN254, BatchContext, BbsError, Ciphersuite, IssuerSet, ProofBytes, PublicKey, SecretKey,
                  ST_ACCEPT, ST_REJECT, ST_ERR_MSG_GEN_LEN, ST_ERR_DISCLOSED_INDEX, ST_ERR_IDX_MSG_LEN,
                  ST_ERR_MALFORMED)

__all__ = ["BLS12_381", "BN254", "BatchContext", "BbsError", "Ciphersuite", "IssuerSet", "ProofBytes", "PublicKey", "SecretKey",
           "ST_ACCEPT", "ST_REJECT", "ST_ERR_MSG_GEN_LEN", "ST_ERR_DISCLOSED_INDEX", "ST_ERR_IDX_MSG_LEN",
           "ST_ERR_MALFORMED"]
