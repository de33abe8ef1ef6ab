"""Drop-in mirrors of /root/reference/utils/* for the MSMD hot path (CUDA, sm_100a)."""
