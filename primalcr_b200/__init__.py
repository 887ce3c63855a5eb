"""primalcr_b200 -- B200-native Primal-CR / Primal-CR++ collaborative-ranking trainer (drop-in for pcr()/pcrpp())."""
__version__ = "0.1.0"
