"""OpenQASM 2.0 front end: text -> a circuit built with the script API.

Same entry points and behaviour as the reference's importer (qgate/openqasm/importer.py:9-50):

    translate(qasm) / translate_file(name)            -> Python source that builds the circuit
    load_circuit(qasm) / load_circuit_from_file(name) -> a module-like object whose attributes are
                                                          the qreg / creg names of the program and
                                                          `circuit` (the op list for Simulator.run)

The reference parses with PLY (lex.py / yacc.py) and exec()s the generated source; PLY is not in this
image, so this is an own hand-written tokenizer + recursive-descent parser of the same grammar
subset (yacc.py:15-232), and `load_circuit` builds the objects directly instead of exec()ing text.
What it accepts, and what each statement becomes, follows the reference's formatter
(formatter.py:67-103):

    U(t, p, l) a            U3(t, p, l)            CX a, b        ctrl(a).X(b)
    u3 / u2 / u1            U3 / U2 / U1           id             I
    x y z h s t             X Y Z H S T            sdg / tdg      S.Adj / T.Adj
    rx / ry                 Rx / Ry                rz             U1   (formatter.py:92-93)
    cx cy cz ch             ctrl(a).X/Y/Z/H(b)     cu1 / cu3      ctrl(a).U1 / U3(b)
    ccx a, b, c             ctrl(a, b).X(c)        swap a, b      Swap(a, b)
    measure q -> c          measure(c, q)          reset q        reset(q)
    barrier ...             barrier(...)           if (c == n) op if_(c, n, [op])

A whole register as an argument broadcasts (`h q;`, `measure q -> c;`, `cx q, r;` pairwise,
`cx q[0], r;` one against all).  Error behaviour is the reference's (tests/test_openqasm.py):
`gate` definitions raise RuntimeError, `opaque` raises NotImplementedError, malformed text raises
SyntaxError with the line number, an undeclared register raises NameError.

`script` selects the namespace the ops are built with: this package's `qgate_b200.script` by default,
or the reference's `qgate.script` — the same text then reaches both front ends.
"""
import math
import re

__all__ = ['translate', 'translate_file', 'load_circuit', 'load_circuit_from_file', 'QasmModule']

_TOKEN = re.compile(r'''
    (?P<ws>\s+|//[^\n]*)
  | (?P<real>(?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)
  | (?P<int>\d+)
  | (?P<str>"[^"\n]*")
  | (?P<id>[a-z][A-Za-z0-9_]*|U\b|CX\b|OPENQASM\b)
  | (?P<sym>->|==|[\[\]\(\)\{\};,+\-*/^])
''', re.VERBOSE)

_KEYWORDS = {'OPENQASM', 'include', 'qreg', 'creg', 'gate', 'opaque', 'measure', 'reset', 'barrier', 'if',
             'pi', 'U', 'CX'}
_FUNCS = {'sin': math.sin, 'cos': math.cos, 'tan': math.tan, 'exp': math.exp, 'ln': math.log,
          'sqrt': math.sqrt}

# qelib1 name -> (builder name, number of parameters, number of controls, adjoint)
_GATES = {
    'u3': ('U3', 3, 0, False), 'u2': ('U2', 2, 0, False), 'u1': ('U1', 1, 0, False),
    'x': ('X', 0, 0, False), 'y': ('Y', 0, 0, False), 'z': ('Z', 0, 0, False),
    'h': ('H', 0, 0, False), 's': ('S', 0, 0, False), 't': ('T', 0, 0, False),
    'sdg': ('S', 0, 0, True), 'tdg': ('T', 0, 0, True), 'id': ('I', 0, 0, False),
    'rx': ('Rx', 1, 0, False), 'ry': ('Ry', 1, 0, False), 'rz': ('U1', 1, 0, False),
    'cx': ('X', 0, 1, False), 'cy': ('Y', 0, 1, False), 'cz': ('Z', 0, 1, False),
    'ch': ('H', 0, 1, False), 'cu1': ('U1', 1, 1, False), 'cu3': ('U3', 3, 1, False),
    'ccx': ('X', 0, 2, False),
}


class _Tok:
    __slots__ = ('kind', 'text', 'line')

    def __init__(self, kind, text, line):
        self.kind, self.text, self.line = kind, text, line


def _tokenize(text):
    pos, line, out = 0, 1, []
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if m is None:
            bad = text[pos:].split(None, 1)[0]
            raise SyntaxError("line {}: illegal character(s) '{}'".format(line, bad[:20]))
        kind = m.lastgroup
        if kind != 'ws':
            out.append(_Tok(kind, m.group(), line))
        line += m.group().count('\n')
        pos = m.end()
    out.append(_Tok('eof', '', line))
    return out


class _Parser:
    """statements as tuples: ('qreg'|'creg', name, n), ('gate', builder, params, adjoint, [args]),
    ('swap', a, b), ('measure', q, c), ('reset', q), ('barrier', [args]), ('if', creg, n, stmt);
    an argument is (name, index or None)."""

    def __init__(self, text):
        self.toks = _tokenize(text)
        self.i = 0

    # -- token helpers
    def peek(self):
        return self.toks[self.i]

    def next(self):
        tok = self.toks[self.i]
        self.i += 1
        return tok

    def error(self, tok, what):
        found = "end of input" if tok.kind == 'eof' else "'{}'".format(tok.text)
        raise SyntaxError('line {}: {}, found {}'.format(tok.line, what, found))

    def expect(self, text):
        tok = self.next()
        if tok.text != text or tok.kind in ('str',):
            self.error(tok, "expected '{}'".format(text))
        return tok

    def accept(self, text):
        if self.peek().text == text and self.peek().kind != 'str':
            self.i += 1
            return True
        return False

    def ident(self):
        tok = self.next()
        if tok.kind != 'id' or tok.text in _KEYWORDS:
            self.error(tok, 'expected an identifier')
        return tok.text

    def integer(self):
        tok = self.next()
        if tok.kind != 'int':
            self.error(tok, 'expected a non-negative integer')
        return int(tok.text)

    # -- grammar (yacc.py:15-232)
    def program(self):
        stmts = []
        if self.peek().text == 'OPENQASM':
            self.next()
            tok = self.next()
            if tok.kind not in ('real', 'int'):
                self.error(tok, 'expected the version number')
            self.expect(';')
        while self.peek().text == 'include':
            self.next()
            tok = self.next()
            if tok.kind != 'str':
                self.error(tok, 'expected a file name')
            if tok.text.strip('"') != 'qelib1.inc':      # yacc.py:47-53: only the standard header
                raise NotImplementedError('line {}: include {} is not supported (qelib1.inc is built in)'
                                          .format(tok.line, tok.text))
            self.expect(';')
        while self.peek().kind != 'eof':
            stmts.append(self.statement())
        return stmts

    def statement(self):
        tok = self.peek()
        if tok.text in ('qreg', 'creg'):
            self.next()
            name = self.ident()
            self.expect('[')
            count = self.integer()
            self.expect(']')
            self.expect(';')
            return (tok.text, name, count)
        if tok.text == 'gate':
            raise RuntimeError('line {}: gate definitions are not supported.'.format(tok.line))
        if tok.text == 'opaque':
            raise NotImplementedError('line {}: opaque gates are not supported.'.format(tok.line))
        if tok.text == 'if':
            self.next()
            self.expect('(')
            creg = self.ident()
            self.expect('==')
            value = self.integer()
            self.expect(')')
            return ('if', creg, value, self.qop())
        if tok.text == 'barrier':
            self.next()
            args = self.arglist()
            self.expect(';')
            return ('barrier', args)
        return self.qop()

    def qop(self):
        tok = self.peek()
        if tok.text == 'measure':
            self.next()
            q = self.argument()
            self.expect('->')
            c = self.argument()
            self.expect(';')
            return ('measure', q, c)
        if tok.text == 'reset':
            self.next()
            q = self.argument()
            self.expect(';')
            return ('reset', q)
        if tok.text == 'U':
            self.next()
            self.expect('(')
            params = self.explist()
            self.expect(')')
            arg = self.argument()
            self.expect(';')
            if len(params) != 3:
                raise SyntaxError('line {}: U takes 3 parameters'.format(tok.line))
            return ('gate', 'U3', params, False, 0, [arg], tok.line)
        if tok.text == 'CX':
            self.next()
            a = self.argument()
            self.expect(',')
            b = self.argument()
            self.expect(';')
            return ('gate', 'X', [], False, 1, [a, b], tok.line)
        if tok.kind != 'id' or tok.text in _KEYWORDS:
            self.error(tok, 'expected a statement')
        name = self.next().text
        params = []
        if self.accept('('):
            if not self.accept(')'):
                params = self.explist()
                self.expect(')')
        args = self.arglist()
        self.expect(';')
        if name == 'swap':
            if params or len(args) != 2:
                raise SyntaxError('line {}: swap takes two qubits'.format(tok.line))
            return ('swap', args[0], args[1], tok.line)
        if name not in _GATES:
            raise RuntimeError("line {}: unknown gate, {}.".format(tok.line, name))   # formatter.py:103
        builder, n_params, n_ctrl, adjoint = _GATES[name]
        if len(params) != n_params:
            raise SyntaxError('line {}: {} takes {} parameter(s)'.format(tok.line, name, n_params))
        if n_ctrl and len(args) != n_ctrl + 1:
            raise SyntaxError('line {}: {} takes {} qubits'.format(tok.line, name, n_ctrl + 1))
        return ('gate', builder, params, adjoint, n_ctrl, args, tok.line)

    def arglist(self):
        args = [self.argument()]
        while self.accept(','):
            args.append(self.argument())
        return args

    def argument(self):
        name = self.ident()
        if self.accept('['):
            index = self.integer()
            self.expect(']')
            return (name, index)
        return (name, None)

    def explist(self):
        exps = [self.exp()]
        while self.accept(','):
            exps.append(self.exp())
        return exps

    # exp : term (('+' | '-') term)* ; term : factor (('*' | '/') factor)* ; factor : ['-'] power ;
    # power : atom ['^' factor] ; atom : REAL | INT | pi | func '(' exp ')' | '(' exp ')'.  An expression is
    # kept as (value, source text) so that translate() can print it the way it was written.
    def exp(self):
        value, src = self.term()
        while self.peek().text in ('+', '-') and self.peek().kind == 'sym':
            op = self.next().text
            rhs, rsrc = self.term()
            value = value + rhs if op == '+' else value - rhs
            src = '{} {} {}'.format(src, op, rsrc)
        return value, src

    def term(self):
        value, src = self.factor()
        while self.peek().text in ('*', '/') and self.peek().kind == 'sym':
            tok = self.next()
            rhs, rsrc = self.factor()
            if tok.text == '/' and rhs == 0.:
                raise SyntaxError('line {}: division by zero'.format(tok.line))
            value = value * rhs if tok.text == '*' else value / rhs
            src = '{} {} {}'.format(src, tok.text, rsrc)
        return value, src

    def factor(self):
        if self.accept('-'):
            value, src = self.factor()
            return -value, '-' + src
        if self.accept('+'):
            return self.factor()
        return self.power()

    def power(self):
        value, src = self.atom()
        if self.accept('^'):
            rhs, rsrc = self.factor()
            return float(value) ** rhs, '{} ** {}'.format(src, rsrc)
        return value, src

    def atom(self):
        tok = self.next()
        if tok.kind in ('real', 'int'):
            return float(tok.text), tok.text
        if tok.text == 'pi':
            return math.pi, 'math.pi'
        if tok.kind == 'id' and tok.text in _FUNCS:
            self.expect('(')
            value, src = self.exp()
            self.expect(')')
            name = 'log' if tok.text == 'ln' else tok.text
            return _FUNCS[tok.text](value), 'math.{}({})'.format(name, src)
        if tok.text == '(' and tok.kind == 'sym':
            value, src = self.exp()
            self.expect(')')
            return value, '({})'.format(src)
        self.error(tok, 'expected an expression')


def _parse(qasm):
    return _Parser(qasm).program()


# ---- direct construction --------------------------------------------------------------------

class QasmModule:
    """what load_circuit returns (importer.py:25-26): attributes = the program's register names and
    `circuit`."""
    pass


class _Builder:
    def __init__(self, script):
        if script is None:
            from . import script
        self.S = script
        self.regs = {}
        self.kinds = {}
        self.circuit = []

    def reg(self, arg, kind, line=None):
        name, index = arg
        if name not in self.regs:
            raise NameError("name '{}' is not defined".format(name))
        if self.kinds[name] != kind:
            raise NameError("'{}' is not a {}".format(name, kind))
        items = self.regs[name]
        if index is None:
            return list(items), True
        if index >= len(items):
            raise IndexError('list index out of range, {}[{}]'.format(name, index))
        return [items[index]], False

    def broadcast(self, operands):
        """[(items, whole register?)] -> tuples, one per application (a whole register against single
        qubits repeats the single ones; several whole registers walk pairwise)."""
        lengths = set(len(items) for items, whole in operands if whole)
        if len(lengths) > 1:
            raise RuntimeError('registers of different sizes in one statement.')
        count = lengths.pop() if lengths else 1
        return [tuple(items[k] if whole else items[0] for items, whole in operands) for k in range(count)]

    def gate_factory(self, builder, params, adjoint, ctrls):
        S = self.S
        holder = S.ctrl(*ctrls) if ctrls else S
        factory = getattr(holder, builder)
        if params:
            factory = factory(*[value for value, _ in params])
        return factory.Adj if adjoint else factory

    def statement(self, stmt, out):
        S, kind = self.S, stmt[0]
        if kind in ('qreg', 'creg'):
            _, name, count = stmt
            self.regs[name] = S.new_qregs(count) if kind == 'qreg' else S.new_references(count)
            self.kinds[name] = kind
        elif kind == 'gate':
            _, builder, params, adjoint, n_ctrl, args, _line = stmt
            if n_ctrl == 0:
                # every argument gets the gate, a whole register one per element (formatter.py:118-130)
                for arg in args:
                    for q in self.reg(arg, 'qreg')[0]:
                        out.append(self.gate_factory(builder, params, adjoint, ())(q))
            else:
                for qs in self.broadcast([self.reg(arg, 'qreg') for arg in args]):
                    if len(set(id(q) for q in qs)) != len(qs):
                        raise RuntimeError('control and target overlap.')
                    out.append(self.gate_factory(builder, params, adjoint, qs[:-1])(qs[-1]))
        elif kind == 'swap':
            for a, b in self.broadcast([self.reg(stmt[1], 'qreg'), self.reg(stmt[2], 'qreg')]):
                out.append(S.Swap(a, b))
        elif kind == 'measure':
            for q, c in self.broadcast([self.reg(stmt[1], 'qreg'), self.reg(stmt[2], 'creg')]):
                out.append(S.measure(c, q))
        elif kind == 'reset':
            for q in self.reg(stmt[1], 'qreg')[0]:
                out.append(S.reset(q))
        elif kind == 'barrier':
            for arg in stmt[1]:
                for q in self.reg(arg, 'qreg')[0]:
                    out.append(S.barrier(q))
        elif kind == 'if':
            _, creg, value, inner = stmt
            refs, _whole = self.reg((creg, None), 'creg')
            clause = []
            self.statement(inner, clause)
            out.append(S.if_(refs, value, clause))
        else:
            raise AssertionError(kind)


def load_circuit(qasm, script=None):
    """OpenQASM 2.0 text -> QasmModule (register names as attributes + `circuit`)."""
    builder = _Builder(script)
    for stmt in _parse(qasm):
        builder.statement(stmt, builder.circuit)
    module = QasmModule()
    for name, items in builder.regs.items():
        setattr(module, name, items)
    module.circuit = builder.circuit
    return module


def load_circuit_from_file(filename, script=None):
    with open(filename, 'r') as file:
        return load_circuit(file.read(), script)


# ---- translation to Python source -------------------------------------------------------------

def _ref(arg):
    name, index = arg
    return name if index is None else '{}[{}]'.format(name, index)


def _emit(stmt, lines, indent, sizes):
    pad = '    ' * indent
    kind = stmt[0]

    def each(arg, template):
        """one line per argument: a comprehension over a whole register"""
        if arg[1] is None:
            lines.append('{}[{} for _{} in {}],'.format(pad, template.format('_' + arg[0]), arg[0], arg[0]))
        else:
            lines.append('{}{},'.format(pad, template.format(_ref(arg))))

    if kind == 'gate':
        _, builder, params, adjoint, n_ctrl, args, _line = stmt
        call = builder + ('({})'.format(', '.join(src for _, src in params)) if params else '') + \
            ('.Adj' if adjoint else '')
        if n_ctrl == 0:
            for arg in args:
                each(arg, call + '({})')
        else:
            whole = [a for a in args if a[1] is None]
            if not whole:
                lines.append('{}ctrl({}).{}({}),'.format(pad, ', '.join(_ref(a) for a in args[:-1]), call,
                                                          _ref(args[-1])))
            else:
                names = [('{}[_k]'.format(a[0]) if a[1] is None else _ref(a)) for a in args]
                lines.append('{}[ctrl({}).{}({}) for _k in range(len({}))],'.format(
                    pad, ', '.join(names[:-1]), call, names[-1], whole[0][0]))
    elif kind == 'swap':
        a, b = stmt[1], stmt[2]
        if a[1] is None or b[1] is None:
            names = [('{}[_k]'.format(x[0]) if x[1] is None else _ref(x)) for x in (a, b)]
            size_of = a[0] if a[1] is None else b[0]
            lines.append('{}[Swap({}, {}) for _k in range(len({}))],'.format(pad, names[0], names[1], size_of))
        else:
            lines.append('{}Swap({}, {}),'.format(pad, _ref(a), _ref(b)))
    elif kind == 'measure':
        q, c = stmt[1], stmt[2]
        if q[1] is None and c[1] is None:
            lines.append('{}[measure(_{c}, _{q}) for _{c}, _{q} in zip({c}, {q})],'.format(pad, c=c[0], q=q[0]))
        else:
            lines.append('{}measure({}, {}),'.format(pad, _ref(c), _ref(q)))
    elif kind == 'reset':
        each(stmt[1], 'reset({})')
    elif kind == 'barrier':
        for arg in stmt[1]:
            each(arg, 'barrier({})')
    elif kind == 'if':
        _, creg, value, inner = stmt
        lines.append('{}if_({}, {}, ['.format(pad, creg, value))
        _emit(inner, lines, indent + 1, sizes)
        lines.append('{}] ),'.format(pad))


def translate(qasm, package='qgate_b200'):
    """OpenQASM 2.0 text -> Python source in the reference's layout (formatter.py:16-64): register
    declarations as assignments, gates collected into `circuit = [...]` / `circuit += [...]` lists."""
    lines = ['import {}'.format(package), 'from {}.script import *'.format(package), 'import math', '']
    opened, initialized, sizes = False, False, {}

    def close():
        nonlocal opened
        if opened:
            lines.extend([']', ''])
            opened = False

    for stmt in _parse(qasm):
        if stmt[0] in ('qreg', 'creg'):
            close()
            sizes[stmt[1]] = stmt[2]
            lines.append('{} = {}({})'.format(stmt[1], 'new_qregs' if stmt[0] == 'qreg' else 'new_references',
                                              stmt[2]))
            continue
        if not opened:
            lines.extend(['', 'circuit += [' if initialized else 'circuit = ['])
            opened = initialized = True
        _emit(stmt, lines, 1, sizes)
    close()
    return '\n'.join(lines) + '\n'


def translate_file(filename, package='qgate_b200'):
    with open(filename, 'r') as file:
        return translate(file.read(), package)
