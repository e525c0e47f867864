/* kiwi_b200 model image: a flat, position-independent, little-endian snapshot of the read-only arrays the
 * lattice-analysis hot path reads.  Every pointer of the reference's in-memory
 * model is replaced by an index; every array is one section.  The image is what `kiwi_init()` of this
 * library loads, what is broadcast once to every GPU, and what stays resident in HBM.
 *
 * Field provenance (reference file:line):
 *   trie   nodes/keys/diffs/values ... include/kiwi/FrozenTrie.h:77-99 (keys re-sorted ascending per node)
 *   forms                          ... include/kiwi/Form.h:231-257
 *   morphemes / chunks             ... include/kiwi/Form.h:142-198
 *   Knlm arrays                    ... src/Knlm.hpp:28-36, include/kiwi/Knlm.h:17-24
 *   CoNg arrays (model_type cong)  ... src/CoNgramModel.hpp:47-65,89-105, include/kiwi/CoNgramModel.h:18-43
 *   config / tag scorer / specials ... include/kiwi/Kiwi.h:150-167,187,222, include/kiwi/TagUtils.h:8-20
 */
#ifndef KIWI_B200_IMAGE_H
#define KIWI_B200_IMAGE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KB2_IMAGE_MAGIC   0x31474D4932424Bull /* "KB2IMG1" */
#define KB2_IMAGE_VERSION 6u
#define KB2_POSTAG_MAX    64                  /* >= (int)POSTag::max of the reference (Types.h:195-227) */

/* section ids */
enum kb2_section_id {
	KB2_SEC_TRIE_NODES = 0,   /* kb2_trie_node[]                                  */
	KB2_SEC_TRIE_KEYS,        /* uint16_t[]  ascending inside each node           */
	KB2_SEC_TRIE_DIFFS,       /* int32_t[]   child = node + diff                  */
	KB2_SEC_FORMS,            /* kb2_form[]                                       */
	KB2_SEC_FORM_CHARS,       /* uint16_t[]  pool of form strings                 */
	KB2_SEC_FORM_CANDS,       /* uint32_t[]  morpheme ids                         */
	KB2_SEC_MORPHS,           /* kb2_morph[]                                      */
	KB2_SEC_MORPH_CHUNKS,     /* kb2_chunk[]                                      */
	KB2_SEC_KN_NODES,         /* kb2_kn_node[]                                    */
	KB2_SEC_KN_KEYS,          /* uint32_t[]  ascending inside each node           */
	KB2_SEC_KN_VALUES,        /* int32_t[]   >0 child diff, <0 leaf ll bit-cast   */
	KB2_SEC_KN_ROOT,          /* int32_t[htx_vocab]  root direct table            */
	KB2_SEC_KN_HTX,           /* uint32_t[vocab] history transform (may be empty) */
	KB2_SEC_CHR_RUNS,         /* kb2_chr_run[]  code-point attribute runs over 0..0x10FFFF, ascending */
	/* CoNg language model (empty for Knlm images) */
	KB2_SEC_CG_NODES,         /* kb2_cg_node[]  non-leaf nodes of the context trie                 */
	KB2_SEC_CG_KEYS,          /* uint32_t[]  ascending inside each node (VL keys, see cg_key_size) */
	KB2_SEC_CG_VALUES,        /* int32_t[]   >0 child diff, <0 leaf: -contextIdx                   */
	KB2_SEC_CG_ROOT,          /* int32_t[cg_root_size]  root direct table (allRootValueData)       */
	KB2_SEC_CG_CTX_EMB,       /* rows: uint8[dim] (= s8 + 128), float scale, float bias            */
	KB2_SEC_CG_OUT_EMB,       /* rows: int8[dim], float scale, int32 hsum (= 128 * sum)            */
	KB2_SEC_CG_INV_VOCAB,     /* uint32_t[vocab] invertedContextVocab (may be empty)               */
	KB2_SEC_CG_OUT_BIAS,      /* float[vocab] outputEmbBias (may be empty)                         */
	/* SkipBigram model on top of the Knlm sections (empty otherwise), src/SkipBigramModel.hpp:24-32 */
	KB2_SEC_SB_PTRS,          /* uint32_t[vocab + 1]  key range of every target token              */
	KB2_SEC_SB_KEYS,          /* uint32_t[]  history tokens, ascending inside each target          */
	KB2_SEC_SB_COMPS,         /* float[]     compensation(target | history), parallel to SB_KEYS   */
	KB2_SEC_SB_DISCNTS,       /* float[vocab]                                                      */
	KB2_SEC_SB_VALID,         /* uint8_t[vocab] vocabValidness                                     */
	KB2_SEC_COUNT
};

typedef struct kb2_section { uint64_t offset, nbytes; } kb2_section;

/* values[] of the form trie: >=0 form index, or one of */
#define KB2_TRIE_NONE      (-1)
#define KB2_TRIE_SUBMATCH  (-2)

typedef struct __attribute__((aligned(16))) kb2_trie_node {
	uint32_t next_offset;   /* into TRIE_KEYS / TRIE_DIFFS                       */
	int32_t  fail;          /* `lower`: relative index of the fail node, 0 = none */
	int32_t  value;         /* form index, KB2_TRIE_NONE or KB2_TRIE_SUBMATCH     */
	uint16_t num_nexts;
	uint16_t depth;
} kb2_trie_node;            /* 16 B: one vector load per visit */

#define KB2_FORM_ZCODA   1u
#define KB2_FORM_ZSIOT   2u
#define KB2_FORM_HASJ    4u
#define KB2_FORM_HASFULL 8u

typedef struct kb2_form {
	uint32_t str_off;       /* into FORM_CHARS */
	uint32_t cand_off;      /* into FORM_CANDS */
	uint16_t str_len;
	uint16_t num_spaces;
	uint16_t cand_cnt;
	uint16_t dialect;
	uint8_t  vowel, polar, flags, form_hash;
} kb2_form;                 /* 20 B */

#define KB2_MORPH_COMPLEX 1u
#define KB2_MORPH_SAISIOT 2u

typedef struct kb2_morph {
	int32_t  form_idx;      /* kform as form index, -1 = null                    */
	int32_t  combined;      /* relative index of the combined morpheme           */
	uint32_t chunk_off;     /* into MORPH_CHUNKS                                 */
	uint32_t lm_morpheme_id;
	uint32_t orig_morpheme_id;
	float    user_score;
	uint16_t dialect;
	uint8_t  tag;           /* POSTag incl. irregular bit 0x80                   */
	uint8_t  vowel;         /* CondVowel                                         */
	uint8_t  polar;         /* CondPolarity                                      */
	uint8_t  flags;         /* KB2_MORPH_*                                       */
	uint8_t  sense_id;
	uint8_t  combine_socket;
	uint8_t  chunk_cnt;
	uint8_t  pad[3];
} kb2_morph;                /* 36 B */

typedef struct kb2_chunk { uint32_t morph; uint8_t begin, end; uint16_t pad; } kb2_chunk;

typedef struct kb2_kn_node {
	uint32_t num_nexts;
	int32_t  lower;
	uint32_t next_offset;
	float    ll, gamma;
} kb2_kn_node;              /* 20 B, include/kiwi/Knlm.h:17-24 */

/* CoNg context-trie node after load (src/CoNgramModel.cpp:484-575): `value` = contextIdx of the node,
 * `lower` = relative index of the longest proper suffix node. */
typedef struct kb2_cg_node { int32_t lower; uint32_t value; uint32_t next_offset; uint32_t num_nexts; } kb2_cg_node;

/* One run of consecutive code points sharing all attributes.  Extracted by calling the reference's pure
 * per-code-point functions over the whole code space: identifySpecialChr (src/Utils.cpp:76-190),
 * chr2ScriptType (src/ScriptType.cpp:5-560), isSpace (include/kiwi/Utils.h:295-326),
 * isEmoji(c0, c1) (src/ScriptType.cpp:569-753; EMOJI1: returns 1 for any c1, EMOJI2: returns 2 when c1 is
 * U+FE0F or a skin-tone modifier). */
#define KB2_CHR_SPACE  1u
#define KB2_CHR_EMOJI1 2u
#define KB2_CHR_EMOJI2 4u
typedef struct kb2_chr_run { uint32_t start; uint8_t cls, script, flags, pad; } kb2_chr_run;

typedef struct kb2_config {   /* KiwiConfig defaults, include/kiwi/Kiwi.h:150-167 */
	float    cut_off_threshold, oov_rule_scale, oov_rule_bias, space_penalty, typo_cost_weight;
	uint32_t max_unk_form_size, max_unk_form_size_followed_by_jclass, space_tolerance;
	uint32_t integrate_allomorph;
} kb2_config;

typedef struct kb2_header {
	uint64_t magic;
	uint32_t version;
	uint32_t model_type;          /* (int)ModelType of Types.h:307: knlm 2, sbg 3, cong 4 */
	uint64_t total_bytes;
	kb2_section sec[KB2_SEC_COUNT];
	uint32_t n_trie_nodes, n_trie_edges, n_forms, n_morphs, n_chunks;
	uint32_t default_tag_size;    /* (int)POSTag::p: first built-in z_coda form = forms[default_tag_size-1+...] */
	uint32_t postag_max;          /* (int)POSTag::max                               */
	uint32_t lang_vocab_size;     /* langMdl->vocabSize()                           */
	/* Knlm scalars (src/Knlm.hpp:28-36) */
	uint32_t kn_num_nodes, kn_num_edges, kn_htx_vocab, kn_has_htx, kn_order;
	int32_t  kn_bos_node;
	float    kn_unk_ll;
	uint32_t special_morph_ids[6];   /* Kiwi::specialMorphIds, Kiwi.h:222           */
	float    tag_left_boundary[2][KB2_POSTAG_MAX];  /* already multiplied by weight, TagUtils.h:16-19 */
	kb2_config config;
	uint32_t n_chr_runs;
	uint32_t script_latin, script_variation_selectors;   /* ScriptType enum values (include/kiwi/ScriptType.h) */
	char     model_name[64];
	/* CoNg scalars (include/kiwi/CoNgramModel.h:18-33); all 0 for Knlm images */
	uint32_t cg_num_nodes, cg_num_edges, cg_root_size, cg_dim, cg_context_size;
	uint32_t cg_key_size;         /* 2: 16-bit keys, 3: 16-bit keys with surrogate pairs for ids >= tMax, 4: 32-bit keys */
	uint32_t cg_flags;            /* CoNgramModelHeader::flags */
	uint32_t cg_pad;
	/* SkipBigram scalars (include/kiwi/SkipBigramModel.h:9-13); 0 for other images */
	uint32_t sb_vocab_size, sb_window_size, sb_num_pairs, sb_pad;
} kb2_header;

#ifdef __cplusplus
}
#endif
#endif
