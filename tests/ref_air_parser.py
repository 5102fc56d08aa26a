"""Evaluates the reference's `evaluate()` bodies straight from their Rust source text (where /root/reference exists): a
mechanical translation of each `eval.add_constraint(<expr>)` and `eval.add_to_relation(RelationEntry::new(..))` into Python
integer arithmetic, so that tests/air_model.py (a hand transcription) and through it csrc/host/air.hpp are checked against
the source itself rather than against a reading of it.  Test infrastructure only; reads, never copies."""
import os
import re

P = (1 << 31) - 1
REF = "/root/reference/crates/brainfuck_prover/src/components"
FILES = ["memory/component.rs", "instruction/component.rs", "program/component.rs", "processor/component.rs",
         "processor/instructions/jump/jump_if_not_zero_component.rs", "processor/instructions/jump/jump_if_zero_component.rs",
         "processor/instructions/input_component.rs", "processor/instructions/left_component.rs",
         "processor/instructions/minus_component.rs", "processor/instructions/output_component.rs",
         "processor/instructions/plus_component.rs", "processor/instructions/right_component.rs",
         "processor/instructions/end_of_execution/component.rs"]
OPCODES = {"Right": ">", "Left": "<", "Plus": "+", "Minus": "-", "PutChar": ".", "ReadChar": ",", "JumpIfZero": "[", "JumpIfNotZero": "]"}
RELATIONS = {"memory_lookup_elements": 0, "instruction_lookup_elements": 1, "processor_lookup_elements": 2}


def available():
    return os.path.isdir(REF)


def _expr(rust: str) -> str:
    e = " ".join(rust.split()).rstrip(", ")          # rustfmt leaves a trailing comma in multi-line calls
    e = e.replace(".clone()", "").replace(".into()", "")
    e = re.sub(r"InstructionType::(\w+)\.to_base_field\(\)", lambda m: str(ord(OPCODES[m.group(1)])), e)
    e = re.sub(r"BaseField::from\((\d+)\)", r"\1", e)
    e = re.sub(r"E::EF::from\((\w+)\)", r"\1", e)
    e = e.replace("BaseField::one()", "1").replace("E::F::one()", "1").replace("E::EF::one()", "1")
    assert re.fullmatch(r"[\w\s+\-*()]*", e), e          # identifiers, integers and + - * ( ) only
    return e


def _balanced(src: str, start: int) -> str:
    """text between the parenthesis at src[start] and its match"""
    depth, i = 0, start
    while True:
        depth += (src[i] == "(") - (src[i] == ")")
        if depth == 0:
            return src[start + 1:i]
        i += 1


def parse(comp: int):
    """-> (column names in mask order, [(kind, ...)] statements in source order)"""
    src = open(os.path.join(REF, FILES[comp])).read()
    body = src[src.index("fn evaluate<E: EvalAtRow>"):]
    body = re.sub(r"//[^\n]*", "", body[:body.index("\n    }\n")])
    cols, prog = [], []
    for m in re.finditer(r"let (\w+)\s*=\s*([^;]*);|eval\.add_constraint\(|eval\.add_to_relation\(", body):
        if m.group(0).startswith("let"):
            name, rhs = m.group(1), m.group(2)
            if "next_trace_mask" in rhs:
                cols.append(name)
            elif "get_preprocessed_column" in rhs:
                prog.append(("is_first", name))
            else:
                prog.append(("let", name, _expr(rhs)))
        elif "add_constraint" in m.group(0):
            prog.append(("constraint", _expr(_balanced(body, m.end() - 1))))
        else:
            inner = _balanced(body, m.end() - 1)                      # RelationEntry::new( &self.X, num, &[..], )
            args = _balanced(inner, inner.index("("))
            rel = RELATIONS[re.search(r"&self\.(\w+)", args).group(1)]
            vals = [v.strip() for v in re.search(r"&\[([^\]]*)\]", args).group(1).replace(".clone()", "").split(",") if v.strip()]
            num = args[args.index(",") + 1:args.index("&[")].strip().rstrip(",")
            prog.append(("relation", rel, _expr(num), vals))
    return cols, prog


def evaluate(comp: int, row, is_first: int):
    """-> (values of the add_constraint calls in order, [(numerator, relation, [values]) per add_to_relation])"""
    cols, prog = parse(comp)
    assert len(cols) == len(row), (comp, cols)
    env = {c.lstrip("_"): v for c, v in zip(cols, row)}
    env.update({c: v for c, v in zip(cols, row)})
    out, rels = [], []
    for st in prog:
        if st[0] == "is_first":
            env[st[1]] = is_first
        elif st[0] == "let":
            env[st[1]] = eval(st[2], {"__builtins__": {}}, env) % P
        elif st[0] == "constraint":
            out.append(eval(st[1], {"__builtins__": {}}, env) % P)
        else:
            rels.append((eval(st[2], {"__builtins__": {}}, env) % P, st[1], [env[v] for v in st[3]]))
    return out, rels


# ---- the seven interaction_trace_evaluation functions (table.rs): which fraction each LogUp column accumulates
TABLE_FILES = {0: "memory/table.rs", 1: "instruction/table.rs", 2: "program/table.rs", 3: "processor/table.rs",
               4: "processor/instructions/jump/table.rs", 5: "processor/instructions/jump/table.rs",
               **{k: "processor/instructions/table.rs" for k in range(6, 12)}, 12: "processor/instructions/end_of_execution/table.rs"}
ELEMENT_TYPES = {"MemoryElements": 0, "InstructionElements": 1, "ProcessorElements": 2}


def parse_fractions(comp: int):
    """-> one entry per `col_gen.write_frac(..)`: (numerator at d = 0, numerator at d = 1, dummy-flag column index or None,
    relation, [main column indices])"""
    src = open(os.path.join(REF, TABLE_FILES[comp])).read()
    src = re.sub(r"//[^\n]*", "", src)
    index_of = {}                                                      # "MemoryColumn::Clk" -> 0
    for em in re.finditer(r"impl (\w+Column) \{.*?fn index\(self\) -> usize \{\s*match self \{(.*?)\}", src, flags=re.S):
        for v, i in re.findall(r"Self::(\w+) => (\d+)", em.group(2)):
            index_of[em.group(1) + "::" + v] = int(i)
    start = src.index("pub fn interaction_trace_evaluation(")
    body = src[start:src.index("\n}\n", start)]
    params = dict(re.findall(r"(\w+): &(\w+Elements)", body[:body.index("{")]))
    col_var = {v: index_of[e + "::" + c] for v, e, c in re.findall(r"let (\w+) = &main_trace_eval\[(\w+)::(\w+)\.index\(\)\]\.data;", body)}
    out = []
    for loop in re.split(r"for vec_row in", body)[1:]:
        loop = loop[:loop.index("write_frac")]
        row_var = {v: col_var[c] for v, c in re.findall(r"let (\w+) = (\w+)\[vec_row\];", loop)}
        num = " ".join(re.search(r"let num\s*=\s*([^;]*);", loop).group(1).split())
        flag = re.search(r"PackedSecureField::from\((\w+)(?:\[vec_row\])?\)", num)
        d_idx = None if not flag else (col_var[flag.group(1)] if flag.group(1) in col_var else row_var[flag.group(1)])
        expr = re.sub(r"PackedSecureField::from\([^)]*\)", "d", num).replace("PackedSecureField::one()", "1")
        assert re.fullmatch(r"[d1\s+\-]*", expr), expr
        comb = re.search(r"(\w+)\s*\.combine\(&\[([^\]]*)\]\)", " ".join(loop.split()))
        vals = [row_var[v.strip()] for v in comb.group(2).split(",") if v.strip()]
        out.append((eval(expr, {}, {"d": 0}), eval(expr, {}, {"d": 1}), d_idx, ELEMENT_TYPES[params[comb.group(1)]], vals))
    return out
