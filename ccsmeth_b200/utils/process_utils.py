"""Constants and small host helpers that are part of the call_mods data contract
(reference ccsmeth/utils/process_utils.py:12-85)."""

# base -> embedding row; every IUPAC ambiguity code collapses to 4 (reference process_utils.py:26-29)
base2code_dna = {'A': 0, 'C': 1, 'G': 2, 'T': 3}
for _b in "NWSMKRYBVDHZ":
    base2code_dna[_b] = 4
code2base_dna = {0: 'A', 1: 'C', 2: 'G', 3: 'T', 4: 'N'}

basepairs = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A', 'N': 'N', 'W': 'W', 'S': 'S', 'M': 'K', 'K': 'M',
             'R': 'Y', 'Y': 'R', 'B': 'V', 'V': 'B', 'D': 'H', 'H': 'D', 'Z': 'Z'}

# model constants (reference process_utils.py:64-73)
N_VOCAB = 5
NEMBED_BASE = 8

max_queue_size = 600
nproc_to_call_mods_in_cpu_mode = 2
default_ref_loc = -1


def str2bool(v):
    return str(v).lower() in ("yes", "true", "t", "1")


def complement_seq(base_seq):
    """Reverse complement (the reference names it complement_seq, process_utils.py:103-118)."""
    return "".join(basepairs.get(b, 'N') for b in reversed(base_seq))
