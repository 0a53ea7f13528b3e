"""Golden vectors transcribed by hand from the reference's own tests (floxer 0.2.0).

Every entry cites the reference test it comes from (file:line relative to the reference tree).
Nothing here is read from /root/reference at test time.
"""

# test/alignment_test.cpp:7-30
ALIGNMENT_SMALL = dict(
    reference=[0, 0, 1, 2, 1, 3, 0, 2, 2, 3, 0, 1],
    query=[1, 2, 1, 3, 1, 2, 2],
    max_errors=2, num_errors=1, start=2, cigar="4=1X2=",
)

# test/verification_test.cpp:11-123
VERIFY_REFERENCE = [
    4, 2, 3, 4, 3, 4, 4, 4, 3, 2,
    4, 3, 3, 2, 2, 3, 4, 4, 3, 3,
    4, 3, 2, 2, 1, 4, 3, 3, 4, 2,
    4, 4, 4, 3, 3, 2, 1, 1, 1, 2,
    3, 4, 4, 3, 2, 4, 4, 2, 1, 4,
    4, 3, 4, 4, 4, 4, 3, 3, 2, 1,
    2, 3, 4, 3, 2, 1, 2, 3, 4, 3,
    1, 4, 2, 1, 4, 4, 2, 2, 3, 4,
    3, 3, 2, 1, 4, 4, 1, 1, 1, 2,
    4, 3, 2, 1, 2, 2, 2, 3, 3, 1,
]
VERIFY_QUERY = [
    4, 3, 4, 4, 4, 4, 3, 3, 2, 1, 4,
    2, 3, 4, 3, 2, 1, 2, 3, 4,
    1, 4, 2, 1, 4, 4, 2, 2, 3, 4,
]
VERIFY = dict(
    tree=dict(total_len=30, num_errors=5, leaf_max_errors=1, strategy="bottom_up"),
    anchor=dict(pex_leaf_index=0, reference_id=0, reference_position=50, num_errors=0),
    extra_verification_ratio=0.1, orientation=1,
    expected=dict(cigar="10=1I9=1D10=", num_errors=2, start=50, orientation=1),
    mutations={5: 1, 6: 1, 11: 3, 20: 2},          # :114-118, after which nothing is added
)

# test/verification_test.cpp:126-161
SPAN = dict(
    anchor_position=100_755, node=(0, 500, 999, 30), leaf_from=750, ref_len=1_000_000,
    cases=[(0.0, (100_475, 561, 0)), (0.01, (100_469, 573, 6))],
)

# test/verification_test.cpp:163-261
NODE_REFERENCE = [2] * 10 + [1] * 80 + [2] * 10
NODE_QUERY = [
    1, 1, 1, 3, 1, 1, 1, 1, 1, 1,
    1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
    1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
    1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
    1, 1, 1, 1, 1, 1, 1, 1, 1, 3,
    1, 4, 1, 1, 1, 2, 1, 1, 1, 1,
    1, 1, 1, 3, 1, 1, 1, 4, 1, 1,
    1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
    1, 1, 1, 1, 1,
]
NODE = dict(node_from=40, node_to=84, num_errors=5, span_offset=50, span_length=50,
            expected=dict(num_errors=5, start=50), extra_mismatch=(42, 2))

# test/intervals_test.cpp:5-33
IVL = dict(
    ivl1=(5, 11), ivl2=(15, 21), ivl3=(11, 14), ivl4=(14, 15), ivl5=(0, 100),
    inside_ivl1=(6, 10), overlapping_below_ivl1=(3, 7), containing_ivl1=(3, 14),
    overlapping_below_ivl2=(13, 18), overlapping_above_ivl2=(17, 23),
    between_both=(11, 15), overlapping_both=(8, 16), containing_both=(3, 30),
    below_both=(0, 2), above_both=(22, 24),
)
# relationship codes follow include/intervals.hpp:14-22
REL = dict(completely_above=0, completely_below=1, contains=2, equal=3, inside=4,
           overlapping_or_touching_above=5, overlapping_or_touching_below=6)
# test/intervals_test.cpp:35-65: (a, b, a.relationship_with(b))
IVL_RELATIONS = [
    ("ivl1", "inside_ivl1", "contains"),
    ("ivl1", "overlapping_below_ivl1", "overlapping_or_touching_above"),
    ("ivl1", "containing_ivl1", "inside"),
    ("ivl1", "overlapping_below_ivl2", "completely_below"),
    ("ivl1", "overlapping_above_ivl2", "completely_below"),
    ("ivl1", "between_both", "overlapping_or_touching_below"),
    ("ivl1", "overlapping_both", "overlapping_or_touching_below"),
    ("ivl1", "containing_both", "inside"),
    ("ivl1", "below_both", "completely_above"),
    ("ivl1", "above_both", "completely_below"),
    ("ivl1", "ivl1", "equal"),
    ("ivl2", "inside_ivl1", "completely_above"),
    ("ivl2", "overlapping_below_ivl1", "completely_above"),
    ("ivl2", "containing_ivl1", "completely_above"),
    ("ivl2", "overlapping_below_ivl2", "overlapping_or_touching_above"),
    ("ivl2", "overlapping_above_ivl2", "overlapping_or_touching_below"),
    ("ivl2", "between_both", "overlapping_or_touching_above"),
    ("ivl2", "overlapping_both", "overlapping_or_touching_above"),
    ("ivl2", "containing_both", "inside"),
    ("ivl2", "below_both", "completely_above"),
    ("ivl2", "above_both", "completely_below"),
    ("ivl2", "ivl2", "equal"),
]
# test/intervals_test.cpp:67-89
IVL_TRIM = [((10, 20), 0, (10, 20)), ((10, 20), 1, (11, 19)), ((10, 20), 5, (14, 15)),
            ((10, 20), 10, (10, 11)), ((10, 20), 25, (10, 11))]
IVL_PROBES = ["inside_ivl1", "overlapping_below_ivl1", "containing_ivl1", "overlapping_below_ivl2",
              "overlapping_above_ivl2", "between_both", "overlapping_both", "containing_both",
              "below_both", "above_both"]
# test/intervals_test.cpp:91-157: after inserting ... the probes above answer
IVL_CONTAINS_STEPS = [
    (["ivl1", "ivl2"], [True] + [False] * 9),
    (["ivl3"], [True] + [False] * 9),
    (["ivl4"], [True] + [False] * 9),
    (["ivl5"], [True] * 10),
]

# test/math_test.cpp:15-25
CEIL_DIV = [((100, 8), 13), ((100, 5), 20)]
CEIL_EPS = [(3.0, 3), (500 * 0.01, 5), (100 * 0.07, 7), (123.456, 124)]

# test/pex_test.cpp:7-143: (total_len, errors, leaf_errors, strategy) -> leaves as (from, length, num_errors)
PEX_LEAVES = [
    ((12, 3, 0, "recursive"), [(0, 3, 0), (3, 3, 0), (6, 3, 0), (9, 3, 0)]),
    ((12, 3, 1, "recursive"), [(0, 6, 1), (6, 6, 1)]),
    ((12, 3, 2, "recursive"), [(0, 6, 1), (6, 6, 1)]),
    ((30, 14, 2, "bottom_up"), [(0, 6, 2), (6, 6, 2), (12, 6, 2), (18, 6, 2), (24, 6, 2)]),
]

# test/data/reference.fasta, test/data/queries.fastq
WHOLE_REFERENCES = {
    "ref": "AAAAAAAAAAAAAAAAACCCCCCCCCCCCCCCCCCCGGGGGGGGGGGGGGGGGGTTTTTTTTTTTTTTTTT",
    '*extra_snippet(")': "ACGTACGT",
}
WHOLE_QUERIES = {
    "query1": "AAACCCGGGTTT", "query2": "AAAAAACCCCCC", "query3": "GGGGAAGGGGGG",
    "query4": "TTTTTTTTTTGG", "query5": "AAAAAAAAAAAC", "query6": "ATATATATATAT",
}
# test/floxer_whole_program_via_cli_test.cpp:17-37,103-112
WHOLE_FLAGS = dict(query_errors=2, seed_errors=[0, 1], extra_verification_ratio=2.0, interval_optimization=True)
# test/floxer_whole_program_via_cli_test.cpp:47-93: (query, reverse_strand) -> (pos_min, pos_max, NM, cigar)
WHOLE_EXPECT = {
    ("query2", True): (48, 48, 0, "12="), ("query2", False): (11, 11, 0, "12="),
    ("query3", True): (17, 26, 2, "6=2I4="), ("query3", False): (36, 44, 2, "4=2I6="),
    ("query4", True): (7, 61, 2, "2I10="), ("query4", False): (54, 61, 2, "10=2I"),
    ("query5", True): (53, 53, 0, "12="), ("query5", False): (6, 6, 0, "12="),
}
WHOLE_UNMAPPED = ["query1", "query6"]
