"""Algorithmic-bytes model of the hot path, evaluated on work counters produced by the
instrumented oracle.  Byte costs come from the reference's structs: FrozenTrie node 12 B + value 8 B
(include/kiwi/FrozenTrie.h:77-99), trie key 2 B, edge diff 4 B, form record 16 B, candidate id 4 B, lattice
node 32 B, morpheme record 32 B, Knlm node 20 B + value 4 B (include/kiwi/Knlm.h:17-24), Knlm key = this model's
key width (4 B: vocabulary > 65535), path record 40 B (src/BestPathContainer.hpp:24-34), output token 15 B."""


def lattice_bytes(c: dict) -> float:
    n, m = c["rawUnits"], c["normUnits"]
    return (2 * n + 2 * m + 4 * (n + c["sentences"])
            + c["trieVisits"] * (12 + 8) + 2 * c["trieProbes"] + 4 * c["trieHits"]
            + 16 * c["candForms"] + 32 * c["nodesBuilt"])


def viterbi_bytes(c: dict, kn_key_bytes: int = 4) -> float:
    return (32 * c["nodesFinal"] + 4 * c["candEntries"] + 32 * c["candEvals"]
            + c["lmHops"] * (20 + 4) + kn_key_bytes * c["lmProbes"]
            + 40 * c["pathsWritten"] + 40 * c["pairs"] + 15 * c["tokens"])


def cong_bytes(c: dict, dim: int = 128) -> float:
    """CoNg scorer on top of the lattice/Viterbi record traffic: gathered embedding rows (context row dim + 16 B,
    output row dim + 8 B in the reference, src/CoNgramModel.hpp:89-105; both dim + 8 here) and context-trie hops
    (node 16 B + key/value packet; lmHops / lmProbes count the CoNg trie when the oracle runs a CoNg image)."""
    return c.get("cgRows", 0) * (dim + 12)


def total_bytes(c: dict) -> float:
    return lattice_bytes(c) + viterbi_bytes(c)
