"""Host side of the witness path: the producer of the `RawPreflightTrace` the device witgen consumes.

The reference's segment prover starts from a `PreflightResults` (risc0/circuit/rv32im/src/prove/witgen/mod.rs:55-88):
the executor runs the guest and cuts it into segments (execute/executor.rs), `Segment::preflight` replays one segment
and records, per cycle, what the circuit's step functions need (prove/witgen/preflight.rs): `RawPreflightCycle` (36 B)
and `RawMemoryTransaction` (20 B) arrays (rv32im-sys/src/lib.rs:21-84), plus the injector CSR scatter and the global
vector (witgen/mod.rs:226-380). That Rust cannot be built in this image (no cargo), so this module restates the part of
it a guest needs:

  execute/rv32im.rs      Emulator (RV32IM decode + step)                    -> Machine._exec
  execute/r0vm.rs        Risc0Machine (resume / suspend / ecalls / traps)   -> Machine
  execute/pager.rs       PagedMemory + paging-cycle accounting              -> PagedMemory
  execute/executor.rs    Executor::run (segment split, claims)              -> execute()
  execute/poseidon2.rs, prove/witgen/poseidon2.rs   paging permutation cycles, zcheck -> Poseidon2State / p2_* / Checksum
  execute/sha2.rs, prove/witgen/sha2.rs             SHA-256 compression cycles -> Sha2State / sha2_ecall
  execute/bibc.rs, execute/bigint.rs, prove/witgen/bigint.rs, prove/witgen/byte_poly.rs
                         bigint2 programs (nondet bytecode + verify program) -> BibcProgram, bigint_ecall(_preflight),
                         BytePolyProgram, BigIntAccum (the mix-dependent accum injector, witgen/mod.rs:186-207)
  binfmt/src/image.rs    MemoryImage (sparse Poseidon2 Merkle image)        -> MemoryImage
  prove/witgen/preflight.rs, witgen/mod.rs          Preflight, injector, globals -> preflight(), PreflightResults

Supported guest surface: RV32IM in machine or user mode, ecall terminate / read (fd supplied by a callback) / write /
poseidon2 / sha2 / bigint, user ecall + mret, fence. It is plain Python: a po2 = 20
segment takes tens of seconds - it feeds tests and the benchmark's setup, not the timed region.
"""
import gzip
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
P = 15 * 2**27 + 1

# ---- platform constants (execute/platform.rs) -----------------------------------------------------------------------
WORD_SIZE = 4
PAGE_BYTES = 1024
PAGE_WORDS = PAGE_BYTES // WORD_SIZE
MEMORY_PAGES = (1 << 32) // PAGE_BYTES
MERKLE_TREE_DEPTH = 22
LOOKUP_TABLE_CYCLES = ((1 << 8) + (1 << 16)) // 16
RESERVED_CYCLES = LOOKUP_TABLE_CYCLES + 1
ZERO_PAGE_END_ADDR = 0x0001_0000
USER_START_ADDR = 0x0001_0000
USER_END_ADDR = 0xc000_0000
KERNEL_START_ADDR = 0xc000_0000
KERNEL_END_ADDR = 0xff00_0000
MACHINE_REGS_ADDR = 0xffff_0000
USER_REGS_ADDR = 0xffff_0080
SAFE_WRITE_ADDR = 0xffff_0100
MEPC_ADDR = 0xffff_0200
SUSPEND_PC_ADDR = 0xffff_0210
SUSPEND_MODE_ADDR = 0xffff_0214
GLOBAL_OUTPUT_ADDR = 0xffff_0240
GLOBAL_INPUT_ADDR = 0xffff_0260
ECALL_DISPATCH_ADDR = 0xffff_1000
TRAP_DISPATCH_ADDR = 0xffff_2000
MEMORY_END_WADDR = 0x4000_0000
MERKLE_TREE_START_WADDR = 0x4000_0000
MERKLE_TREE_END_WADDR = 0x4400_0000
POVW_NONCE_START_WADDR = 0x4400_0000
POVW_NONCE_END_WADDR = 0x4400_0008
REG_MAX = 32
REG_A0, REG_A1, REG_A2, REG_A3, REG_A4, REG_A7 = 10, 11, 12, 13, 14, 17
REG_SP, REG_T0, REG_T1, REG_T2, REG_T3 = 2, 5, 6, 7, 28
USER_BIGINT_END_ADDR = 0xbfff_0000
HOST_ECALL_TERMINATE, HOST_ECALL_READ, HOST_ECALL_WRITE, HOST_ECALL_POSEIDON2, HOST_ECALL_SHA2, HOST_ECALL_BIGINT = range(6)
PFLAG_IS_ELEM, PFLAG_CHECK_OUT = 0x8000_0000, 0x4000_0000
MAX_IO_BYTES, MAX_IO_WORDS = 1024, 4
MAX_SHA_COUNT = 10           # platform.rs:137
MAX_INSN_CYCLES, MAX_INSN_CYCLES_LOWER_PO2 = 25_000, 2_000   # rv32im/src/lib.rs:45-48
DIGEST_WORDS = 8


class CS:  # CycleState
    LoadRootAndNonce, Resume, Suspend, StoreRoot, ControlTable, ControlDone = 0, 1, 4, 5, 6, 7
    MachineEcall, Terminate, HostReadSetup, HostWrite, HostReadBytes, HostReadWords = 8, 9, 10, 11, 12, 13
    PoseidonEntry, PoseidonLoadState, PoseidonLoadIn, PoseidonDoOut, PoseidonPaging, PoseidonStoreState = 16, 17, 18, 21, 22, 23
    PoseidonExtRound, PoseidonIntRound = 24, 25
    ShaEcall, ShaLoadState, ShaLoadData, ShaMix, ShaStoreState = 32, 33, 34, 35, 36
    BigIntEcall, BigIntStep, Decode = 40, 41, 48


MAJOR_CONTROL0, MAJOR_ECALL0 = 7, 8
CONTROL_USER_ECALL, CONTROL_FENCE, CONTROL_MRET = 2, 2, 3
TX_READ, TX_PAGE_IN, TX_PAGE_OUT = 0, 1, 2

# InsnKind (execute/rv32im.rs:96-158): major = kind // 8, minor = kind % 8
(ADD, SUB, XOR, OR, AND, SLT, SLTU, ADDI, XORI, ORI, ANDI, SLTI, SLTIU, BEQ, BNE, BLT, BGE, BLTU, BGEU, JAL, JALR, LUI,
 AUIPC) = range(23)
SLL, SLLI, MUL, MULH, MULHSU, MULHU = range(24, 30)
SRL, SRA, SRLI, SRAI, DIV, DIVU, REM, REMU = range(32, 40)
LB, LH, LW, LBU, LHU = range(40, 45)
SB, SH, SW = range(48, 51)
EANY, MRET, FENCE = 56, 57, 58

# pager.rs cycle costs
PAGE_CYCLES = 1 + 10 * (PAGE_WORDS // 8) + 1
NODE_CYCLES = 1 + 2 + 8 + 1 + 1
RESERVED_PAGING_CYCLES = 1 + 1 + 1 + 2 + 2 + 1 + 1 + 1
POSEIDON_PAGE_ROUNDS = PAGE_WORDS // 8
UNLOADED, LOADED, DIRTY = 0, 1, 2

M32 = 0xFFFFFFFF


# ---- Poseidon2 over BabyBear in normal form (zkp/src/core/hash/poseidon2, execute/poseidon2.rs:150-238) ---------------
def _load_p2_tables():
    import re
    text = open(os.path.join(HERE, "csrc", "tables", "poseidon2_tables.h")).read()

    def arr(name):
        m = re.search(r"%s\[\d+\]\s*=\s*\{([^}]*)\}" % name, text)
        return [int(x, 0) for x in re.findall(r"0x[0-9a-fA-F]+|\d+", m.group(1))]

    return arr("R0_P2_RC_FULL"), arr("R0_P2_RC_PARTIAL"), arr("R0_P2_DIAG")


RC_FULL, RC_PARTIAL, M_INT_DIAG = _load_p2_tables()   # 8 x 24 full-round constants, 21 partial, 24 diagonal (normal form)


def _sbox(x):
    x2 = x * x % P
    x4 = x2 * x2 % P
    return x4 * x2 % P * x % P


def p2_m_ext(c):
    out = [0] * 24
    sums = [0, 0, 0, 0]
    for i in range(6):
        x0, x1, x2, x3 = c[4 * i:4 * i + 4]
        t0 = x0 + x1
        t1 = x2 + x3
        t2 = 2 * x1 + t1
        t3 = 2 * x3 + t0
        t4 = 4 * t1 + t3
        t5 = 4 * t0 + t2
        t6 = t3 + t5
        t7 = t2 + t4
        ch = (t6 % P, t5 % P, t7 % P, t4 % P)
        for j in range(4):
            out[4 * i + j] = ch[j]
            sums[j] += ch[j]
    return [(out[i] + sums[i % 4]) % P for i in range(24)]


def p2_ext_round(c, idx):
    """full round idx in 0..7 (add constants, sbox, M_ext)"""
    rc = RC_FULL[24 * idx:24 * idx + 24]
    return p2_m_ext([_sbox((c[i] + rc[i]) % P) for i in range(24)])


def p2_int_rounds(c):
    c = list(c)
    for r in range(21):
        c[0] = _sbox((c[0] + RC_PARTIAL[r]) % P)
        s = sum(c) % P
        c = [(s + M_INT_DIAG[i] * c[i]) % P for i in range(24)]
    return c


def poseidon2_mix(c):
    c = p2_m_ext(c)
    for i in range(4):
        c = p2_ext_round(c, i)
    c = p2_int_rounds(c)
    for i in range(4, 8):
        c = p2_ext_round(c, i)
    return c


# ---- MemoryImage (binfmt/src/image.rs) ---------------------------------------------------------------------------------
def page_digest(words):
    cells = [0] * 24
    for i in range(PAGE_WORDS // DIGEST_WORDS):
        for j in range(DIGEST_WORDS):
            w = words[i * DIGEST_WORDS + j]
            cells[2 * j] = w & 0xffff
            cells[2 * j + 1] = w >> 16
        cells = poseidon2_mix(cells)
    return tuple(cells[:8])


def digest_pair(lhs, rhs):
    cells = list(rhs) + list(lhs) + [0] * 8
    return tuple(poseidon2_mix(cells)[:8])


_ZERO_DIGESTS = None


def zero_digests():
    global _ZERO_DIGESTS
    if _ZERO_DIGESTS is None:
        d = page_digest([0] * PAGE_WORDS)
        out = [None] * (MERKLE_TREE_DEPTH + 1)
        for depth in range(MERKLE_TREE_DEPTH, -1, -1):
            out[depth] = d
            d = digest_pair(d, d)
        _ZERO_DIGESTS = out
    return _ZERO_DIGESTS


class MemoryImage:
    def __init__(self):
        self.pages = {}      # page_idx -> list of 256 words
        self.digests = {1: zero_digests()[0]}
        self.dirty = set()

    @staticmethod
    def from_words(image):
        """image: {byte address: word}"""
        this = MemoryImage()
        pages = {}
        for addr, word in sorted(image.items()):
            waddr = addr // 4
            pages.setdefault(waddr // PAGE_WORDS, [0] * PAGE_WORDS)[waddr % PAGE_WORDS] = word & M32
        for idx in sorted(pages):
            this.set_page(idx, pages[idx])
        this.update_digests()
        return this

    @staticmethod
    def new_kernel(entry, image):
        image = dict(image)
        image[SUSPEND_PC_ADDR] = entry
        image[SUSPEND_MODE_ADDR] = 1
        return MemoryImage.from_words(image)

    @staticmethod
    def new_user(entry, image):
        image = dict(image)
        image[USER_START_ADDR] = entry
        return MemoryImage.from_words(image)

    def clone(self):
        c = MemoryImage()
        c.pages = {k: list(v) for k, v in self.pages.items()}
        c.digests = dict(self.digests)
        c.dirty = set(self.dirty)
        return c

    def get_page_indexes(self):
        return sorted(self.pages)

    def is_zero(self, idx):
        depth = idx.bit_length() - 1
        while idx not in self.digests and idx > 0:
            idx //= 2
            depth -= 1
        return idx != 0 and self.digests[idx] == zero_digests()[depth]

    def expand_if_zero(self, idx):
        if not self.is_zero(idx):
            return False
        depth = idx.bit_length() - 1
        while idx not in self.digests:
            parent = idx // 2
            self.digests[2 * parent] = zero_digests()[depth]
            self.digests[2 * parent + 1] = zero_digests()[depth]
            idx = parent
            depth -= 1
        return True

    def get_page(self, page_idx):
        if page_idx in self.pages:
            return self.pages[page_idx]
        if self.expand_if_zero(MEMORY_PAGES + page_idx):
            self.pages[page_idx] = [0] * PAGE_WORDS
            return self.pages[page_idx]
        raise KeyError("Unavailable page: %d" % page_idx)

    def set_page(self, page_idx, words, digest=None):
        idx = MEMORY_PAGES + page_idx
        self.expand_if_zero(idx)
        self.digests[idx] = digest if digest is not None else page_digest(words)
        self.pages[page_idx] = list(words)
        self.mark_dirty(idx)

    def get_digest(self, idx):
        self.expand_if_zero(idx)
        return self.digests[idx]

    def set_digest(self, idx, digest):
        self.expand_if_zero(idx)
        self.digests[idx] = digest
        self.mark_dirty(idx)

    def image_id(self):
        return self.get_digest(1)

    def mark_dirty(self, idx):
        while idx != 1:
            parent = idx // 2
            if 2 * parent in self.digests and 2 * parent + 1 in self.digests:
                self.dirty.add(parent)
                idx = parent
            else:
                break

    def update_digests(self):
        for idx in sorted(self.dirty, reverse=True):
            self.digests[idx] = digest_pair(self.digests[2 * idx], self.digests[2 * idx + 1])
        self.dirty = set()


def compute_partial_image(input_image, indexes):
    """pager.rs:572-598"""
    image = MemoryImage()
    idxs = set(indexes)
    for node_idx in sorted(i for i in idxs if i >= MEMORY_PAGES):
        page_idx = node_idx - MEMORY_PAGES
        image.set_page(page_idx, input_image.get_page(page_idx), input_image.get_digest(node_idx))
    for node_idx in sorted(i for i in idxs if i < MEMORY_PAGES):
        for child in (2 * node_idx, 2 * node_idx + 1):
            if child not in idxs:
                image.set_digest(child, input_image.get_digest(child))
    image.update_digests()
    return image


# ---- PagedMemory (execute/pager.rs) ------------------------------------------------------------------------------------
class PagedMemory:
    def __init__(self, image):
        self.image_pages = {k: list(v) for k, v in image.pages.items()}   # WorkingImage
        regs_page = image.get_page(MACHINE_REGS_ADDR // 4 // PAGE_WORDS)
        self.machine_registers = [regs_page[(MACHINE_REGS_ADDR // 4 + i) % PAGE_WORDS] for i in range(REG_MAX)]
        self.user_registers = [regs_page[(USER_REGS_ADDR // 4 + i) % PAGE_WORDS] for i in range(REG_MAX)]
        self.cache = {}          # page_idx -> list (loaded pages)
        self.states = {}         # node_idx -> LOADED / DIRTY (insertion ordered)
        self.cycles = RESERVED_PAGING_CYCLES

    def reset(self):
        self.cache = {}
        self.states = {}
        self.cycles = RESERVED_PAGING_CYCLES

    def page_indexes(self):
        return list(self.states.keys())

    def _image_page(self, page_idx):
        if page_idx not in self.image_pages:
            self.image_pages[page_idx] = [0] * PAGE_WORDS
        return self.image_pages[page_idx]

    def _reg(self, waddr):
        u, m = USER_REGS_ADDR // 4, MACHINE_REGS_ADDR // 4
        if u <= waddr < u + REG_MAX:
            return self.user_registers, waddr - u
        if m <= waddr < m + REG_MAX:
            return self.machine_registers, waddr - m
        return None, 0

    def peek(self, waddr):
        if waddr >= MEMORY_END_WADDR:
            raise ValueError("Invalid peek address: %#x" % waddr)
        regs, i = self._reg(waddr)
        if regs is not None:
            return regs[i]
        page_idx = waddr // PAGE_WORDS
        page = self.cache.get(page_idx)
        if page is None:
            page = self._image_page(page_idx)
        return page[waddr % PAGE_WORDS]

    def _fixup_costs(self, node_idx, goal):
        while node_idx != 0:
            state = self.states.get(node_idx, UNLOADED)
            if goal > state:
                if node_idx < MEMORY_PAGES:
                    if state == UNLOADED:
                        self.cycles += NODE_CYCLES
                    if goal == DIRTY:
                        self.cycles += NODE_CYCLES
                self.states[node_idx] = goal
            node_idx //= 2

    def _load_page(self, page_idx):
        self.cache[page_idx] = list(self._image_page(page_idx))
        self.cycles += PAGE_CYCLES
        self._fixup_costs(MEMORY_PAGES + page_idx, LOADED)

    def load(self, waddr):
        if waddr >= MEMORY_END_WADDR:
            raise ValueError("Invalid load address: %#x" % waddr)
        regs, i = self._reg(waddr)
        if regs is not None:
            return regs[i]
        page_idx = waddr // PAGE_WORDS
        page = self.cache.get(page_idx)
        if page is None:
            self._load_page(page_idx)
            self.states.setdefault(MEMORY_PAGES + page_idx, LOADED)
            page = self.cache[page_idx]
        return page[waddr % PAGE_WORDS]

    def _page_for_writing(self, page_idx):
        node_idx = MEMORY_PAGES + page_idx
        state = self.states.get(node_idx, UNLOADED)
        if state == UNLOADED:
            self._load_page(page_idx)
            state = LOADED
        if state == LOADED:
            self.cycles += PAGE_CYCLES
            self._fixup_costs(node_idx, DIRTY)
            self.states[node_idx] = DIRTY
        return self.cache[page_idx]

    def store(self, waddr, word):
        if waddr >= MEMORY_END_WADDR:
            raise ValueError("Invalid store address: %#x" % waddr)
        regs, i = self._reg(waddr)
        if regs is not None:
            regs[i] = word
            return
        self._page_for_writing(waddr // PAGE_WORDS)[waddr % PAGE_WORDS] = word

    def commit(self):
        """write the cached registers back, return {page_idx: words} of the dirty pages (and fold them into the image)"""
        page = self._page_for_writing(MACHINE_REGS_ADDR // 4 // PAGE_WORDS)
        for i in range(REG_MAX):
            page[(MACHINE_REGS_ADDR // 4 + i) % PAGE_WORDS] = self.machine_registers[i]
            page[(USER_REGS_ADDR // 4 + i) % PAGE_WORDS] = self.user_registers[i]
        out = {}
        for node_idx in sorted(self.states):
            if node_idx >= MEMORY_PAGES and self.states[node_idx] == DIRTY:
                page_idx = node_idx - MEMORY_PAGES
                self.image_pages[page_idx] = list(self.cache[page_idx])
                out[page_idx] = list(self.cache[page_idx])
        return out

    def loaded_pages(self):
        return PagingActivity(sorted(self.image_pages))

    def dirty_pages(self):
        return PagingActivity(sorted(n - MEMORY_PAGES for n, s in self.states.items() if n >= MEMORY_PAGES and s == DIRTY))


class PagingActivity:
    """preflight.rs:715-736"""

    def __init__(self, pages):
        self.pages = list(pages)
        nodes = set()
        for page_idx in self.pages:
            node_idx = MEMORY_PAGES + page_idx
            while node_idx != 1:
                parent = node_idx // 2
                if parent in nodes:
                    break
                nodes.add(parent)
                node_idx = parent
        self.nodes = sorted(nodes)


# ---- the machine (execute/rv32im.rs Emulator + execute/r0vm.rs Risc0Machine) ----------------------------------------
PEEK, LOAD, RECORD = 0, 1, 2


def _sx(x):
    return x - (1 << 32) if x & 0x80000000 else x


class Trap(Exception):
    pass


class Machine:
    """ctx must provide: pc, machine_mode, load_u32(op, waddr), store_u32(waddr, word), on_insn_end(kind),
    on_ecall_cycle(cur, nxt, s0, s1, s2), on_terminate(a0, a1), host_read(fd, n) -> bytes, host_write(fd, bytes) -> rlen,
    on_poseidon2_cycle(cur_state, p2), trap_rewind(), suspend_hook(), resume_hook()"""

    def __init__(self, ctx):
        self.c = ctx

    # Risc0Machine::resume / suspend (r0vm.rs:331-347)
    def resume(self):
        c = self.c
        pc = c.load_u32(RECORD, SUSPEND_PC_ADDR // 4)
        if pc < ZERO_PAGE_END_ADDR:
            raise ValueError("%#x is an invalid guest address" % pc)
        mode = c.load_u32(RECORD, SUSPEND_MODE_ADDR // 4)
        c.pc, c.machine_mode = pc, mode
        c.resume_hook()

    def suspend(self):
        c = self.c
        c.store_u32(SUSPEND_PC_ADDR // 4, c.pc)
        c.store_u32(SUSPEND_MODE_ADDR // 4, c.machine_mode)
        c.suspend_hook()

    def _regs_base(self):
        return (MACHINE_REGS_ADDR if self.c.machine_mode != 0 else USER_REGS_ADDR) // 4

    def load_register(self, idx):
        return self.c.load_u32(RECORD, self._regs_base() + idx)

    def store_register(self, idx, word):
        base = self._regs_base()
        self.c.store_u32(base + (REG_MAX * 2 if idx == 0 else idx), word & M32)

    def _check_insn_load(self, addr):
        return not (addr < ZERO_PAGE_END_ADDR or (self.c.machine_mode == 0 and addr >= KERNEL_START_ADDR))

    def _check_data(self, addr):
        return (addr >= ZERO_PAGE_END_ADDR and self.c.machine_mode != 0) or (USER_START_ADDR <= addr < USER_END_ADDR)

    def _enter_trap(self, dispatch_addr):
        c = self.c
        if c.machine_mode != 0:
            raise Trap("Illegal trap in machine mode")
        c.store_u32(MEPC_ADDR // 4, c.pc)
        c.pc = dispatch_addr
        c.machine_mode = 1

    def trap(self, cause):
        c = self.c
        c.trap_rewind()
        dispatch = c.load_u32(RECORD, TRAP_DISPATCH_ADDR // 4 + cause)
        if dispatch & 3 or not (KERNEL_START_ADDR <= dispatch < KERNEL_END_ADDR):
            raise Trap("Invalid trap address: %#x, cause: %d" % (dispatch, cause))
        self._enter_trap(dispatch)
        return False

    def step(self):
        c = self.c
        pc = c.pc
        if not self._check_insn_load(pc):
            self.trap(1)
            return
        word = c.load_u32(RECORD, pc // 4)
        if word & 3 != 3:
            self.trap(2)
            return
        kind = self._exec(word)
        if kind is not None:
            c.on_insn_end(kind)

    def _exec(self, insn):
        c = self.c
        opcode, rd, f3 = insn & 0x7f, (insn >> 7) & 31, (insn >> 12) & 7
        rs1, rs2, f7 = (insn >> 15) & 31, (insn >> 20) & 31, insn >> 25
        top = insn >> 31
        imm_i = (top * 0xfffff000) | (f7 << 5) | rs2
        kind = None
        if opcode == 0b0110011:
            kind = {(0, 0): ADD, (0, 32): SUB, (1, 0): SLL, (2, 0): SLT, (3, 0): SLTU, (5, 0): SRL, (4, 0): XOR, (5, 32): SRA,
                    (6, 0): OR, (7, 0): AND, (0, 1): MUL, (1, 1): MULH, (2, 1): MULHSU, (3, 1): MULHU, (4, 1): DIV,
                    (5, 1): DIVU, (6, 1): REM, (7, 1): REMU}.get((f3, f7))
        elif opcode == 0b0010011:
            kind = {0: ADDI, 2: SLTI, 3: SLTIU, 4: XORI, 6: ORI, 7: ANDI}.get(f3)
            if f3 == 1 and f7 == 0:
                kind = SLLI
            elif f3 == 5 and f7 == 0:
                kind = SRLI
            elif f3 == 5 and f7 == 32:
                kind = SRAI
        elif opcode == 0b0000011:
            kind = {0: LB, 1: LH, 2: LW, 4: LBU, 5: LHU}.get(f3)
            if kind is not None:
                return self._load(kind, rd, rs1, imm_i)
        elif opcode == 0b0100011:
            kind = {0: SB, 1: SH, 2: SW}.get(f3)
            if kind is not None:
                return self._store(kind, rs1, rs2, (top * 0xfffff000) | (f7 << 5) | rd)
        elif opcode == 0b0110111:
            kind = LUI
        elif opcode == 0b0010111:
            kind = AUIPC
        elif opcode == 0b1100011:
            kind = {0: BEQ, 1: BNE, 4: BLT, 5: BGE, 6: BLTU, 7: BGEU}.get(f3)
        elif opcode == 0b1101111:
            kind = JAL
        elif opcode == 0b1100111:
            kind = JALR
        elif opcode == 0b1110011 and f3 == 0 and f7 == 0b0011000:
            return self._system(MRET, rs2, insn)
        elif opcode == 0b1110011 and f3 == 0 and f7 == 0:
            return self._system(EANY, rs2, insn)
        elif opcode == 0b0001111 and f3 == 0:
            return self._system(FENCE, rs2, insn)
        if kind is None:
            return FENCE + 197 if self.trap(2) else None   # unreachable: trap() returns False
        # step_compute (rv32im.rs:349-458)
        pc = c.pc
        new_pc = (pc + 4) & M32
        a = self.load_register(rs1)
        b = a if rs1 == rs2 else self.load_register(rs2)
        out = 0
        if kind == ADD:
            out = a + b
        elif kind == SUB:
            out = a - b
        elif kind == XOR:
            out = a ^ b
        elif kind == OR:
            out = a | b
        elif kind == AND:
            out = a & b
        elif kind == SLL:
            out = a << (b & 31)
        elif kind == SRL:
            out = a >> (b & 31)
        elif kind == SRA:
            out = _sx(a) >> (b & 31)
        elif kind == SLT:
            out = int(_sx(a) < _sx(b))
        elif kind == SLTU:
            out = int(a < b)
        elif kind == ADDI:
            out = a + imm_i
        elif kind == XORI:
            out = a ^ imm_i
        elif kind == ORI:
            out = a | imm_i
        elif kind == ANDI:
            out = a & imm_i
        elif kind == SLLI:
            out = a << (imm_i & 31)
        elif kind == SRLI:
            out = a >> (imm_i & 31)
        elif kind == SRAI:
            out = _sx(a) >> (imm_i & 31)
        elif kind == SLTI:
            out = int(_sx(a) < _sx(imm_i))
        elif kind == SLTIU:
            out = int(a < imm_i)
        elif BEQ <= kind <= BGEU:
            cond = {BEQ: a == b, BNE: a != b, BLT: _sx(a) < _sx(b), BGE: _sx(a) >= _sx(b), BLTU: a < b, BGEU: a >= b}[kind]
            rd = 0
            if cond:
                imm_b = (top * 0xfffff000) | ((rd_field(insn) & 1) << 11) | ((f7 & 0x3f) << 5) | (rd_field(insn) & 0x1e)
                new_pc = (pc + imm_b) & M32
        elif kind == JAL:
            imm_j = (top * 0xfff00000) | (rs1 << 15) | (f3 << 12) | ((rs2 & 1) << 11) | ((f7 & 0x3f) << 5) | (rs2 & 0x1e)
            new_pc = (pc + imm_j) & M32
            out = pc + 4
        elif kind == JALR:
            new_pc = (a + imm_i) & 0xfffffffe
            out = pc + 4
        elif kind == LUI:
            out = insn & 0xfffff000
        elif kind == AUIPC:
            out = pc + (insn & 0xfffff000)
        elif kind == MUL:
            out = a * b
        elif kind == MULH:
            out = (_sx(a) * _sx(b)) >> 32
        elif kind == MULHSU:
            out = (_sx(a) * b) >> 32
        elif kind == MULHU:
            out = (a * b) >> 32
        elif kind == DIV:
            out = M32 if b == 0 else _trunc_div(_sx(a), _sx(b))
        elif kind == DIVU:
            out = M32 if b == 0 else a // b
        elif kind == REM:
            out = a if b == 0 else _trunc_rem(_sx(a), _sx(b))
        elif kind == REMU:
            out = a if b == 0 else a % b
        if new_pc & 3:
            return kind if self.trap(0) else None
        self.store_register(rd, out & M32)
        c.pc = new_pc
        return kind

    def _load(self, kind, rd, rs1, imm_i):
        c = self.c
        a = self.load_register(rs1)
        addr = (a + imm_i) & M32
        if not self._check_data(addr):
            return kind if self.trap(5) else None
        data = c.load_u32(RECORD, addr // 4)
        shift = 8 * (addr & 3)
        if kind == LB:
            out = (data >> shift) & 0xff
            if out & 0x80:
                out |= 0xffffff00
        elif kind == LH:
            if addr & 1:
                return kind if self.trap(4) else None
            out = (data >> shift) & 0xffff
            if out & 0x8000:
                out |= 0xffff0000
        elif kind == LW:
            if addr & 3:
                return kind if self.trap(4) else None
            out = data
        elif kind == LBU:
            out = (data >> shift) & 0xff
        else:
            if addr & 1:
                return kind if self.trap(4) else None
            out = (data >> shift) & 0xffff
        self.store_register(rd, out)
        c.pc = (c.pc + 4) & M32
        return kind

    def _store(self, kind, rs1, rs2, imm_s):
        c = self.c
        a = self.load_register(rs1)
        b = a if rs1 == rs2 else self.load_register(rs2)
        addr = (a + imm_s) & M32
        shift = 8 * (addr & 3)
        if not self._check_data(addr):
            return kind if self.trap(7) else None
        data = c.load_u32(RECORD, addr // 4)
        if kind == SB:
            data ^= data & (0xff << shift)
            data |= (b & 0xff) << shift
        elif kind == SH:
            if addr & 1:
                return kind if self.trap(6) else None
            data ^= data & (0xffff << shift)
            data |= (b & 0xffff) << shift
        else:
            if addr & 3:
                return kind if self.trap(6) else None
            data = b
        c.store_u32(addr // 4, data & M32)
        c.pc = (c.pc + 4) & M32
        return kind

    def _system(self, kind, rs2, insn):
        c = self.c
        if kind == EANY:
            if rs2 == 0:
                ok = self._machine_ecall() if c.machine_mode != 0 else self._user_ecall()
            elif rs2 == 1:
                ok = self.trap(3)
            else:
                ok = self.trap(2)
            return kind if ok else None
        if kind == MRET:
            if c.machine_mode == 0:
                raise Trap("Illegal mret in user mode")
            dispatch = c.load_u32(RECORD, MEPC_ADDR // 4)
            if dispatch < ZERO_PAGE_END_ADDR:
                raise ValueError("invalid guest address")
            c.pc = (dispatch + 4) & M32
            c.machine_mode = 0
            return kind
        c.pc = (c.pc + 4) & M32   # fence
        return kind

    def _user_ecall(self):
        c = self.c
        dispatch = c.load_u32(RECORD, ECALL_DISPATCH_ADDR // 4)
        if dispatch < ZERO_PAGE_END_ADDR:
            raise ValueError("invalid guest address")
        if dispatch & 3 or not (KERNEL_START_ADDR <= dispatch < KERNEL_END_ADDR):
            return self.trap(9)
        self._enter_trap(dispatch)
        return True

    def _machine_ecall(self):
        which = self.load_register(REG_A7)
        c = self.c
        if which == HOST_ECALL_TERMINATE:
            c.on_ecall_cycle(CS.MachineEcall, CS.Terminate, 0, 0, 0)
            a0 = self.load_register(REG_A0)
            a1 = self.load_register(REG_A1)
            c.on_terminate(a0, a1)
            c.pc = (c.pc + 4) & M32
            c.on_ecall_cycle(CS.Terminate, CS.Suspend, 0, 0, 0)
            return False
        if which == HOST_ECALL_READ:
            return self._ecall_read()
        if which == HOST_ECALL_WRITE:
            c.on_ecall_cycle(CS.MachineEcall, CS.HostWrite, 0, 0, 0)
            fd = self.load_register(REG_A0)
            ptr = self.load_register(REG_A1)
            ln = self.load_register(REG_A2)
            if ptr + ln > M32 + 1 or ln > MAX_IO_BYTES:
                raise ValueError("Invalid length in host write: %d" % ln)
            data = bytes(_peek_u8(c, ptr + i) for i in range(ln))
            rlen = c.host_write(fd, data)
            self.store_register(REG_A0, rlen)
            c.pc = (c.pc + 4) & M32
            c.on_ecall_cycle(CS.HostWrite, CS.Decode, 0, 0, 0)
            return False
        if which == HOST_ECALL_POSEIDON2:
            c.pc = (c.pc + 4) & M32
            c.on_ecall_cycle(CS.MachineEcall, CS.PoseidonEntry, 0, 0, 0)
            m = MACHINE_REGS_ADDR // 4
            # Address convention: this snapshot's Rust (execute/poseidon2.rs:285-293, zkvm/platform syscall.rs:482-505)
            # hands the three registers over as WORD addresses, but the circuit it ships (steps.cpp:6059-6069 ReadAddr:
            # high * 2^14 + low / 4) takes them as BYTE addresses and divides by four - a trace built the Rust way is
            # rejected by the reference's own compiled witgen ("Inconsistent set" at stateAddr / bufOutAddr). The
            # circuit is what the witness generator implements, so the restatement follows the circuit.
            regs = [c.load_u32(RECORD, m + r) for r in (REG_A0, REG_A1, REG_A2, REG_A3)]
            if any(r % 4 for r in regs[:3]):
                raise ValueError("poseidon2 ecall: unaligned address")
            p2 = Poseidon2State.new_ecall(regs[0] // 4, regs[1] // 4, regs[2] // 4, regs[3])
            p2.rest(c, CS.Decode)
            return False
        if which == HOST_ECALL_SHA2:          # r0vm.rs:559-571
            c.pc = (c.pc + 4) & M32
            c.on_ecall_cycle(CS.MachineEcall, CS.ShaEcall, 0, 0, 0)
            sha2_ecall(c)
            return False
        if which == HOST_ECALL_BIGINT:        # r0vm.rs:573-585
            c.pc = (c.pc + 4) & M32
            c.on_ecall_cycle(CS.MachineEcall, CS.BigIntEcall, 0, 0, 0)
            c.ecall_bigint()
            return False
        raise ValueError("unknown machine ecall %d" % which)

    def _ecall_read(self):
        c = self.c
        c.on_ecall_cycle(CS.MachineEcall, CS.HostReadSetup, 0, 0, 0)
        cur = [CS.HostReadSetup]
        fd = self.load_register(REG_A0)
        ptr = self.load_register(REG_A1)
        ln = self.load_register(REG_A2)
        if ptr + ln > M32 + 1:
            raise ValueError("Invalid length in host read: %d" % ln)
        if ln > MAX_IO_BYTES:
            raise ValueError("Invalid length (too big) in host read: %d" % ln)
        if ln > 0 and ptr < ZERO_PAGE_END_ADDR:
            raise ValueError("invalid guest address")
        data = c.host_read(fd, ln)
        rlen = len(data)
        self.store_register(REG_A0, rlen)
        if rlen == 0:
            c.pc = (c.pc + 4) & M32

        def add_cycle(p, r):
            if r == 0:
                nxt = CS.Decode
            elif p & 3 or r < WORD_SIZE:
                nxt = CS.HostReadBytes
            else:
                nxt = CS.HostReadWords
            c.on_ecall_cycle(cur[0], nxt, p // 4, p & 3, r)
            cur[0] = nxt

        def store_u8(addr, byte):
            word = c.load_u32(RECORD, addr // 4)
            sh = 8 * (addr & 3)
            c.store_u32(addr // 4, (word & ~(0xff << sh) & M32) | (byte << sh))

        add_cycle(ptr, rlen)
        i = 0
        while rlen > 0 and ptr & 3:
            store_u8(ptr, data[i])
            ptr, i, rlen = ptr + 1, i + 1, rlen - 1
            if rlen == 0:
                c.pc = (c.pc + 4) & M32
            add_cycle(ptr, rlen)
        while rlen >= MAX_IO_WORDS:
            words = min(rlen // MAX_IO_WORDS, MAX_IO_WORDS)
            for j in range(MAX_IO_WORDS):
                if j < words:
                    c.store_u32(ptr // 4, int.from_bytes(data[i:i + 4], "little"))
                    ptr, i, rlen = ptr + 4, i + 4, rlen - 4
                else:
                    c.store_u32(SAFE_WRITE_ADDR // 4 + j, 0)
            if rlen == 0:
                c.pc = (c.pc + 4) & M32
            add_cycle(ptr, rlen)
        while rlen > 0:
            store_u8(ptr, data[i])
            ptr, i, rlen = ptr + 1, i + 1, rlen - 1
            if rlen == 0:
                c.pc = (c.pc + 4) & M32
            add_cycle(ptr, rlen)
        return False


def rd_field(insn):
    return (insn >> 7) & 31


def _trunc_div(a, b):
    q = abs(a) // abs(b)
    return -q if (a < 0) != (b < 0) else q


def _trunc_rem(a, b):
    r = abs(a) % abs(b)
    return -r if a < 0 else r


def _peek_u8(c, addr):
    return (c.load_u32(PEEK, addr // 4) >> (8 * (addr & 3))) & 0xff


# ---- SHA-256 compression cycles (execute/sha2.rs, prove/witgen/sha2.rs) -----------------------------------------------
SHA2_LOAD_STATE_CYCLES, SHA2_LOAD_DATA_CYCLES, SHA2_MIX_CYCLES, SHA2_STORE_CYCLES = 4, 16, 48, 4
SHA2_BACK = SHA2_LOAD_STATE_CYCLES + SHA2_LOAD_DATA_CYCLES + SHA2_MIX_CYCLES


def _bswap(x):
    return int.from_bytes(int(x).to_bytes(4, "little"), "big")


def _rotr(x, n):
    return ((x >> n) | (x << (32 - n))) & M32


class Sha2State:
    """execute/sha2.rs:31-43: what a SHA cycle's row carries (7 field columns + the bits of a, e, w)"""
    FIELDS = ("state_in_addr", "state_out_addr", "data_addr", "count", "k_addr", "round", "next_state", "a", "e", "w")

    def __init__(self, **kw):
        for f in self.FIELDS:
            setattr(self, f, kw.get(f, 0))

    def clone(self):
        return Sha2State(**{f: getattr(self, f) for f in self.FIELDS})

    def fp_array(self):
        return [self.state_in_addr, self.state_out_addr, self.data_addr, self.count, self.k_addr, self.round, self.next_state]

    def u32_array(self):
        return [self.a, self.e, self.w]


class _Ring:
    def __init__(self, n):
        self.buf, self.cur, self.n = [0] * n, 0, n

    def push(self, v):
        self.buf[self.cur] = v
        self.cur = (self.cur + 1) % self.n

    def back(self, i):
        return self.buf[(self.n + self.cur - i) % self.n]


def _sha_compute_ae(old_a, old_e, k, w):
    a, b, c, d = (old_a.back(i) for i in (1, 2, 3, 4))
    e, f, g, h = (old_e.back(i) for i in (1, 2, 3, 4))
    t1 = (h + (_rotr(e, 6) ^ _rotr(e, 11) ^ _rotr(e, 25)) + ((e & f) ^ (~e & M32 & g)) + k + w) & M32
    t2 = ((_rotr(a, 2) ^ _rotr(a, 13) ^ _rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c))) & M32
    return (t1 + t2) & M32, (d + t1) & M32


def _sha_compute_w(old_w):
    x2, x15 = old_w.back(2), old_w.back(15)
    s1 = _rotr(x2, 17) ^ _rotr(x2, 19) ^ (x2 >> 10)
    s0 = _rotr(x15, 7) ^ _rotr(x15, 18) ^ (x15 >> 3)
    return (s1 + old_w.back(7) + s0 + old_w.back(16)) & M32


def _guest_waddr(addr):
    if addr < ZERO_PAGE_END_ADDR:      # r0vm.rs:733-740 guest_addr
        raise ValueError("%#010x is an invalid guest address" % addr)
    return addr // 4


def sha2_ecall(c):
    """execute/sha2.rs:58-150: `count` SHA-256 compressions of 16-word blocks at a2 from the state at a0 into a1, round
    constants read from the guest's own table at a4; one cycle per state word pair / data word / mix round."""
    m = MACHINE_REGS_ADDR // 4
    sha = Sha2State(state_in_addr=_guest_waddr(c.load_u32(RECORD, m + REG_A0)),
                    state_out_addr=_guest_waddr(c.load_u32(RECORD, m + REG_A1)),
                    data_addr=_guest_waddr(c.load_u32(RECORD, m + REG_A2)),
                    count=c.load_u32(RECORD, m + REG_A3) & 0xffff,
                    k_addr=_guest_waddr(c.load_u32(RECORD, m + REG_A4)), next_state=CS.ShaEcall)
    if sha.count > MAX_SHA_COUNT:
        raise ValueError("Invalid count (too big) in sha2 ecall: %d" % sha.count)
    cur = [CS.ShaEcall]

    def step(nxt):
        sha.next_state = nxt
        c.on_sha2_cycle(cur[0], sha)
        cur[0] = nxt

    old_a, old_e, old_w = _Ring(SHA2_BACK), _Ring(SHA2_BACK), _Ring(16)
    for i in range(SHA2_LOAD_STATE_CYCLES):
        sha.round = i
        step(CS.ShaLoadState)
        a = c.load_u32(RECORD, sha.state_in_addr + 3 - i)
        e = c.load_u32(RECORD, sha.state_in_addr + 7 - i)
        sha.a, sha.e = _bswap(a), _bswap(e)
        old_a.push(sha.a)
        old_e.push(sha.e)
        c.store_u32(sha.state_out_addr + 3 - i, a)
        c.store_u32(sha.state_out_addr + 7 - i, e)
    while sha.count != 0:
        for i in range(SHA2_LOAD_DATA_CYCLES):
            sha.round = i
            step(CS.ShaLoadData)
            k = c.load_u32(RECORD, sha.k_addr + i)
            sha.w = _bswap(c.load_u32(RECORD, sha.data_addr))
            sha.data_addr += 1
            old_w.push(sha.w)
            sha.a, sha.e = _sha_compute_ae(old_a, old_e, k, sha.w)
            old_a.push(sha.a)
            old_e.push(sha.e)
        for i in range(SHA2_MIX_CYCLES):
            sha.round = i
            step(CS.ShaMix)
            k = c.load_u32(RECORD, sha.k_addr + 16 + i)
            sha.w = _sha_compute_w(old_w)
            old_w.push(sha.w)
            sha.a, sha.e = _sha_compute_ae(old_a, old_e, k, sha.w)
            old_a.push(sha.a)
            old_e.push(sha.e)
        for i in range(SHA2_STORE_CYCLES):
            sha.round = i
            step(CS.ShaStoreState)
            sha.a = (old_a.back(4) + old_a.back(SHA2_BACK)) & M32
            sha.e = (old_e.back(4) + old_e.back(SHA2_BACK)) & M32
            sha.w = 0
            if i == 3:
                sha.count -= 1
            old_a.push(sha.a)
            old_e.push(sha.e)
            c.store_u32(sha.state_out_addr + 3 - i, _bswap(sha.a))
            c.store_u32(sha.state_out_addr + 7 - i, _bswap(sha.e))
    sha.round = 0
    step(CS.Decode)


# ---- BigInt ecall (execute/bibc.rs, execute/bigint.rs, prove/witgen/bigint.rs, prove/witgen/byte_poly.rs) ------------
BIGINT_WIDTH_WORDS, BIGINT_WIDTH_BYTES = 4, 16
(BI_CONST, BI_LOAD, BI_STORE, BI_ADD, BI_SUB, BI_MUL, BI_REM, BI_QUO, BI_INV) = (0x2, 0x3, 0x4, 0x8, 0x9, 0xA, 0xB, 0xC, 0xE)
POLY_RESET, POLY_SHIFT, POLY_SET_TERM, POLY_ADD_TOTAL, POLY_CARRY1, POLY_CARRY2, POLY_EQ_ZERO = range(7)
MEM_READ, MEM_WRITE, MEM_CHECK = range(3)


class BibcProgram:
    """execute/bibc.rs:117-160: the 'nondet' half of a bigint blob - inputs, types, 64-bit constants, 64-bit ops"""

    def __init__(self, data):
        import struct
        assert data[:4] == b"bibc" and struct.unpack_from("<I", data, 4)[0] == 1, "Invalid BigInt2 bytecode"
        ni, nt, nc, no = struct.unpack_from("<4I", data, 8)
        p = 24 + 16 * ni                                             # inputs are not needed to evaluate
        self.types = [struct.unpack_from("<4Q", data, p + 32 * i)[0] for i in range(nt)]   # coeffs
        p += 32 * nt
        self.constants = list(struct.unpack_from("<%dQ" % nc, data, p))
        p += 8 * nc
        self.ops = []
        for bits in struct.unpack_from("<%dQ" % no, data, p):
            code = bits & 0xF
            if code not in (BI_CONST, BI_LOAD, BI_STORE, BI_ADD, BI_SUB, BI_MUL, BI_REM, BI_QUO, BI_INV):
                raise ValueError("Invalid BigInt2 bytecode")
            self.ops.append((code, (bits >> 4) & 0xFFF, (bits >> 16) & 0xFFFFFF, (bits >> 40) & 0xFFFFFF))

    def eval(self, io):
        """bibc.rs:162-221 over Python integers (`/` and `%` of malachite's Integer truncate towards zero)"""
        regs = [0] * len(self.ops)
        for i, (code, rt, a, b) in enumerate(self.ops):
            if code == BI_CONST:
                regs[i] = sum(self.constants[a + k] << (64 * k) for k in range(b))
            elif code == BI_LOAD:
                regs[i] = io.load(a >> 16, a & 0xFFFF, -(-self.types[rt] // 16) * 16)
            elif code == BI_STORE:
                io.store(a >> 16, a & 0xFFFF, -(-self.types[rt] // 16) * 16, abs(regs[b]))
            else:
                assert a < i and b < i
                x, y = regs[a], regs[b]
                if code == BI_ADD:
                    regs[i] = x + y
                elif code == BI_SUB:
                    regs[i] = x - y
                elif code == BI_MUL:
                    regs[i] = x * y
                elif code == BI_REM:
                    regs[i] = abs(x) % abs(y) * (-1 if x < 0 else 1)
                elif code == BI_QUO:
                    regs[i] = abs(x) // abs(y) * (-1 if (x < 0) != (y < 0) else 1)
                else:
                    regs[i] = pow(abs(x) % abs(y), -1, abs(y))     # raises ValueError when not invertible


def _aligned_waddr(addr):
    if addr % 4:
        raise ValueError("%#010x is an unaligned address" % addr)
    return addr // 4


def _check_bigint_addr(waddr, mode):
    if not ((waddr >= ZERO_PAGE_END_ADDR // 4 and mode == 1) or waddr < USER_BIGINT_END_ADDR // 4):
        raise ValueError("Invalid bigint address")


class _BigIntIO:
    """execute/bigint.rs:37-133: loads read guest memory (no transactions), stores only fill the witness map"""

    def __init__(self, c, mode):
        self.c, self.mode, self.witness = c, mode, {}

    def _base(self, arena):
        return _aligned_waddr(self.c.load_u32(LOAD, MACHINE_REGS_ADDR // 4 + arena))

    def load(self, arena, offset, count):
        start = self._base(arena) + offset * BIGINT_WIDTH_WORDS
        _check_bigint_addr(start, self.mode)
        words = -(-count // 4)
        limbs = [self.c.load_u32(LOAD, start + i) for i in range(words)]
        if limbs and count % 4:
            limbs[-1] &= (1 << (8 * (count % 4))) - 1
        return sum(w << (32 * i) for i, w in enumerate(limbs))

    def store(self, arena, offset, count, value):
        addr = self._base(arena) + offset * BIGINT_WIDTH_WORDS
        _check_bigint_addr(addr, self.mode)
        nlimbs = (value.bit_length() + 31) // 32
        if count < nlimbs * 4:
            raise ValueError("bigint_store: count (%d bytes) too small for value (%d bytes)" % (count, nlimbs * 4))
        if count % BIGINT_WIDTH_BYTES:
            raise ValueError("bigint_store: count (%d) is not a multiple of %d" % (count, BIGINT_WIDTH_BYTES))
        raw = value.to_bytes(nlimbs * 4, "little")
        filled = 0
        for ci in range(0, len(raw), BIGINT_WIDTH_BYTES):
            chunk = raw[ci:ci + BIGINT_WIDTH_BYTES]
            self.witness[addr + (ci // BIGINT_WIDTH_BYTES) * BIGINT_WIDTH_WORDS] = chunk + b"\x00" * (BIGINT_WIDTH_BYTES - len(chunk))
            filled += 1
        for i in range(count // BIGINT_WIDTH_BYTES - filled):
            self.witness[addr + (filled + i) * BIGINT_WIDTH_WORDS] = b"\x00" * BIGINT_WIDTH_BYTES


def bigint_ecall(c):
    """execute/bigint.rs:172-226 -> (mode, verify_program_ptr (word address of the word BEFORE the program),
    verify_program_size, witness {word address: 16 bytes}). Only the loads of t0 and t2 are recorded transactions."""
    m = MACHINE_REGS_ADDR // 4
    mode = c.load_u32(RECORD, m + REG_T0)
    if mode not in (0, 1):
        raise ValueError("Invalid mode for bigint ecall: %d" % mode)
    blob_ptr = _aligned_waddr(c.load_u32(LOAD, m + REG_A0))
    nondet_ptr = _aligned_waddr(c.load_u32(LOAD, m + REG_T1))
    verify_ptr = _aligned_waddr(c.load_u32(RECORD, m + REG_T2)) - 1
    consts_ptr = _aligned_waddr(c.load_u32(LOAD, m + REG_T3))
    nondet_size = c.load_u32(LOAD, blob_ptr)
    verify_size = c.load_u32(LOAD, blob_ptr + 1)
    consts_size = c.load_u32(LOAD, blob_ptr + 2)
    data = b"".join(c.load_u32(LOAD, nondet_ptr + i).to_bytes(4, "little") for i in range(nondet_size))
    io = _BigIntIO(c, mode)
    BibcProgram(data).eval(io)
    for i in range(verify_size):
        c.load_u32(LOAD, verify_ptr + 1 + i)
    for i in range(consts_size):
        c.load_u32(LOAD, consts_ptr + i)
    return mode, verify_ptr, verify_size, io.witness


# polynomials with small integer coefficients, lowest degree first; lengths follow the reference's 4-lane chunks
# (byte_poly.rs:128-262) because the carry pass walks `len()` coefficients and `get` is bounds-checked
def _bp_add(a, b):
    n = max(len(a), len(b))
    return [(a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0) for i in range(n)]


def _bp_mul(a, b):
    r = [0] * (len(a) + len(b))
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                r[i + j] += x * y
    return r


class BytePolyProgram:
    """byte_poly.rs:33-126: what the verify program computes, over the integers"""

    def __init__(self):
        self.in_carry = False
        self.total_carry = []
        self.reset()

    def reset(self):
        self.poly, self.term, self.total = [0] * 4, [1, 0, 0, 0], [0] * 4

    def step(self, poly_op, coeff, witness):
        delta = list(witness)
        new_poly = _bp_add(self.poly, delta)
        if poly_op == POLY_RESET:
            self.reset()
        elif poly_op == POLY_SHIFT:
            self.poly = [0] * BIGINT_WIDTH_BYTES + new_poly
        elif poly_op == POLY_SET_TERM:
            self.poly, self.term = [0] * 4, new_poly
        elif poly_op == POLY_ADD_TOTAL:
            self.total = _bp_add(self.total, [x * coeff for x in _bp_mul(new_poly, self.term)])
            self.term, self.poly = [1, 0, 0, 0], [0] * 4
        elif poly_op == POLY_CARRY1:
            self.poly = _bp_add(self.poly, [(x - 128) * 64 * 256 for x in delta])
        elif poly_op == POLY_CARRY2:
            self.poly = _bp_add(self.poly, [x * 256 for x in delta])
        elif poly_op == POLY_EQ_ZERO:
            self.total = _bp_add(self.total, _bp_mul([-256, 1, 0, 0], new_poly))
            if any(self.total):
                raise ValueError("Invalid eqz in bigint program")
            self.reset()
            self.in_carry = False
        else:
            raise ValueError("Invalid poly_op in bigint program")


class BigIntState:
    """prove/witgen/bigint.rs:36-45: the 22 injected cells of a bigint cycle"""

    def __init__(self, is_ecall, mode, pc, poly_op, coeff, bytes_, next_state):
        self.is_ecall, self.mode, self.pc, self.poly_op, self.coeff = is_ecall, mode, pc, poly_op, coeff
        self.bytes, self.next_state = bytes(bytes_), next_state

    def clone(self):
        return BigIntState(self.is_ecall, self.mode, self.pc, self.poly_op, self.coeff, self.bytes, self.next_state)

    def as_array(self):
        return [int(self.is_ecall), self.mode, self.pc, self.poly_op, self.coeff] + list(self.bytes) + [self.next_state]


def bigint_ecall_preflight(c):
    """prove/witgen/bigint.rs:104-190,248-268: one BigIntEcall cycle, then one BigIntStep cycle per word of the verify
    program - each reads / writes / checks one 16-byte chunk and advances the byte polynomials"""
    mode, verify_ptr, _size, witness = bigint_ecall(c)
    st = BigIntState(True, mode, verify_ptr, POLY_RESET, 0, bytes(16), CS.BigIntStep)
    prog = BytePolyProgram()
    c.on_bigint_cycle(CS.BigIntEcall, st)
    m = MACHINE_REGS_ADDR // 4
    while st.next_state == CS.BigIntStep:
        st.pc += 1
        insn = c.load_u32(RECORD, st.pc)
        mem_op, poly_op = (insn >> 28) & 0xF, (insn >> 24) & 0xF
        if mem_op > MEM_CHECK:
            raise ValueError("Invalid mem_op in bigint program")
        if poly_op > POLY_EQ_ZERO:
            raise ValueError("Invalid poly_op in bigint program")
        coeff, reg, offset = ((insn >> 21) & 7) - 4, (insn >> 16) & 0x1F, insn & 0xFFFF
        addr = _aligned_waddr(c.load_u32(RECORD, m + reg)) + offset * BIGINT_WIDTH_WORDS
        if mem_op == MEM_CHECK and poly_op != POLY_RESET:
            if not prog.in_carry:
                prog.in_carry = True
                tc, carry = list(prog.total), 0
                for i in range(len(tc)):
                    v = tc[i] + carry
                    if v % 256:
                        raise ValueError("bad carry")
                    tc[i] = carry = v // 256
                prog.total_carry = tc
            out = bytearray(16)
            for i in range(16):
                value = (prog.total_carry[offset * BIGINT_WIDTH_BYTES + i] + 128 * 256 * 64) & M32
                if poly_op == POLY_CARRY1:
                    out[i] = (value >> 14) & 0xFF
                elif poly_op == POLY_CARRY2:
                    out[i] = (value >> 8) & 0x3F
                elif poly_op in (POLY_SHIFT, POLY_EQ_ZERO):
                    out[i] = value & 0xFF
                else:
                    raise ValueError("Invalid poly_op in bigint program")
            st.bytes = bytes(out)
        elif mem_op == MEM_READ:
            st.bytes = b"".join(c.load_u32(RECORD, addr + i).to_bytes(4, "little") for i in range(BIGINT_WIDTH_WORDS))
        elif addr != 0:
            if addr not in witness:
                raise ValueError("Missing bigint witness: %#x" % (addr * 4))
            st.bytes = witness[addr]
            if mem_op == MEM_WRITE:
                for i in range(BIGINT_WIDTH_WORDS):
                    c.store_u32(addr + i, int.from_bytes(st.bytes[4 * i:4 * i + 4], "little"))
        prog.step(poly_op, coeff, st.bytes)
        st.is_ecall = False
        st.poly_op, st.coeff = poly_op, coeff + 4
        st.next_state = CS.Decode if poly_op == POLY_RESET else CS.BigIntStep
        c.on_bigint_cycle(CS.BigIntStep, st)


def ext_sub(a, b):
    return tuple((x - y) % P for x, y in zip(a, b))


class BigIntAccum:
    """byte_poly.rs:403-475: the same program evaluated at the accum mix point (an extension element) - the values of the
    BigIntAccumState registers, which depend on the transcript's mix and are therefore injected in the accum phase
    (witgen/mod.rs:186-207)"""

    def __init__(self, mix):
        self.powers, cur = [], (1, 0, 0, 0)
        for _ in range(BIGINT_WIDTH_BYTES + 1):
            self.powers.append(cur)
            cur = ext_mul(cur, mix)
        self.neg_poly = (0, 0, 0, 0)
        for pw in self.powers[:BIGINT_WIDTH_BYTES]:
            self.neg_poly = ext_add(self.neg_poly, ext_mul(pw, (128, 0, 0, 0)))
        self.reset()

    def reset(self):
        self.poly, self.term, self.total = (0, 0, 0, 0), (1, 0, 0, 0), (0, 0, 0, 0)

    def step(self, st):
        delta = (0, 0, 0, 0)
        for b, pw in zip(st.bytes, self.powers):
            delta = ext_add(delta, ext_mul(pw, (b, 0, 0, 0)))
        new_poly = ext_add(self.poly, delta)
        op = st.poly_op
        if op == POLY_RESET:
            self.reset()
        elif op == POLY_SHIFT:
            self.poly = ext_mul(new_poly, self.powers[BIGINT_WIDTH_BYTES])
        elif op == POLY_SET_TERM:
            self.poly, self.term = (0, 0, 0, 0), new_poly
        elif op == POLY_ADD_TOTAL:
            coeff = ((st.coeff - 4) % P, 0, 0, 0)
            self.total = ext_add(self.total, ext_mul(ext_mul(coeff, self.term), new_poly))
            self.poly, self.term = (0, 0, 0, 0), (1, 0, 0, 0)
        elif op == POLY_CARRY1:
            self.poly = ext_add(self.poly, ext_mul(ext_sub(delta, self.neg_poly), (64 * 256, 0, 0, 0)))
        elif op == POLY_CARRY2:
            self.poly = ext_add(self.poly, ext_mul(delta, (256, 0, 0, 0)))
        elif op == POLY_EQ_ZERO:
            goal = ext_add(self.total, ext_mul(new_poly, ext_sub(self.powers[1], (256, 0, 0, 0))))
            if any(goal):
                raise ValueError("Invalid eqz in bigint accum")
            self.reset()

    def as_array(self):
        return list(self.poly) + list(self.term) + list(self.total)


# ---- Poseidon2 cycles (execute/poseidon2.rs:37-148, prove/witgen/poseidon2.rs) --------------------------------------
class Poseidon2State:
    FIELDS = ("has_state", "state_addr", "buf_out_addr", "is_elem", "check_out", "load_tx_type", "next_state", "sub_state",
              "buf_in_addr", "count", "mode")

    def __init__(self, **kw):
        for f in self.FIELDS:
            setattr(self, f, 0)
        self.next_state = CS.LoadRootAndNonce
        self.inner = [0] * 24
        self.zcheck = (0, 0, 0, 0)
        for k, v in kw.items():
            setattr(self, k, v)

    def clone(self):
        c = Poseidon2State()
        for f in self.FIELDS:
            setattr(c, f, getattr(self, f))
        c.inner = list(self.inner)
        c.zcheck = self.zcheck
        return c

    @staticmethod
    def new_ecall(state_addr, buf_in_addr, buf_out_addr, bits_count):
        return Poseidon2State(state_addr=state_addr, buf_in_addr=buf_in_addr, buf_out_addr=buf_out_addr,
                              has_state=int(state_addr != 0), is_elem=int(bits_count & PFLAG_IS_ELEM != 0),
                              check_out=int(bits_count & PFLAG_CHECK_OUT != 0), count=bits_count & 0xffff, mode=1,
                              load_tx_type=TX_READ, next_state=CS.PoseidonEntry)

    @staticmethod
    def new_start(mode):
        return Poseidon2State(buf_out_addr=MERKLE_TREE_END_WADDR if mode == 0 else MERKLE_TREE_START_WADDR, is_elem=1,
                              check_out=1, load_tx_type=TX_PAGE_IN, next_state=CS.PoseidonPaging, mode=mode)

    @staticmethod
    def new_done(buf_out_addr, next_state, mode):
        return Poseidon2State(buf_out_addr=buf_out_addr, next_state=next_state, mode=mode)

    @staticmethod
    def new_node(node_idx, is_read):
        return Poseidon2State(buf_out_addr=node_idx_to_addr(node_idx), is_elem=1, check_out=int(is_read),
                              load_tx_type=TX_PAGE_IN if is_read else TX_PAGE_OUT, next_state=CS.PoseidonPaging,
                              buf_in_addr=node_idx_to_addr(2 * node_idx + 1), count=1, mode=0 if is_read else 4)

    @staticmethod
    def new_page(page_idx, is_read):
        return Poseidon2State(buf_out_addr=node_idx_to_addr(MEMORY_PAGES + page_idx), check_out=int(is_read),
                              load_tx_type=TX_PAGE_IN if is_read else TX_PAGE_OUT, next_state=CS.PoseidonPaging,
                              buf_in_addr=page_idx * PAGE_WORDS, count=POSEIDON_PAGE_ROUNDS, mode=1 if is_read else 3)

    def _step(self, ctx, cur, nxt, sub):
        self.next_state, self.sub_state = nxt, sub
        ctx.on_poseidon2_cycle(cur[0], self)
        cur[0] = nxt

    def rest(self, ctx, final_state):
        cur = [self.next_state]
        if self.has_state == 1:
            self._step(ctx, cur, CS.PoseidonLoadState, 0)
            for i in range(8):
                self.inner[16 + i] = ctx.load_u32(RECORD, self.state_addr + i)
        addr = self.buf_in_addr
        while self.count > 0:
            self._step(ctx, cur, CS.PoseidonLoadIn, 0)
            if self.is_elem:
                for i in range(8):
                    self.inner[i] = ctx.load_u32(RECORD, addr)
                    addr += 1
                self.buf_in_addr = addr
                self._step(ctx, cur, CS.PoseidonLoadIn, 1)
                for i in range(8):
                    self.inner[8 + i] = ctx.load_u32(RECORD, addr)
                    addr += 1
                self.buf_in_addr = addr
            else:
                for i in range(8):
                    w = ctx.load_u32(RECORD, addr)
                    addr += 1
                    self.inner[2 * i] = w & 0xffff
                    self.inner[2 * i + 1] = w >> 16
                self.buf_in_addr = addr
            self.inner = p2_m_ext(self.inner)
            for i in range(4):
                self._step(ctx, cur, CS.PoseidonExtRound, i)
                self.inner = p2_ext_round(self.inner, i)
            self._step(ctx, cur, CS.PoseidonIntRound, 0)
            self.inner = p2_int_rounds(self.inner)
            for i in range(4, 8):
                self._step(ctx, cur, CS.PoseidonExtRound, i)
                self.inner = p2_ext_round(self.inner, i)
            self.count -= 1
        self._step(ctx, cur, CS.PoseidonDoOut, 0)
        out = self.buf_out_addr
        if self.check_out:
            for i in range(8):
                w = ctx.load_u32(RECORD, out + i)
                if w != self.inner[i]:
                    raise ValueError("poseidon2 check failed: %#010x != %#010x (cell %d, out %#x)" % (w, self.inner[i], i, out))
        else:
            for i in range(8):
                ctx.store_u32(out + i, self.inner[i])
        self.buf_in_addr = 0
        if self.has_state == 1:
            self._step(ctx, cur, CS.PoseidonStoreState, 0)
            for i in range(8):
                ctx.store_u32(self.state_addr + i, self.inner[16 + i])
        self._step(ctx, cur, final_state, 0)

    def as_array(self):
        return ([getattr(self, f) for f in self.FIELDS] + list(self.inner) + [int(x) for x in self.zcheck])


def node_addr_to_idx(waddr):
    return (MERKLE_TREE_END_WADDR - waddr) // DIGEST_WORDS


def node_idx_to_addr(idx):
    return MERKLE_TREE_END_WADDR - idx * DIGEST_WORDS


def get_digest_addr(idx):
    return MERKLE_TREE_START_WADDR + DIGEST_WORDS * (2 * MEMORY_PAGES - idx)


# extension field arithmetic in normal form (for zcheck), X^4 = -11
def ext_mul(a, b):
    nb = P - 11
    return ((a[0] * b[0] + nb * (a[1] * b[3] + a[2] * b[2] + a[3] * b[1])) % P,
            (a[0] * b[1] + a[1] * b[0] + nb * (a[2] * b[3] + a[3] * b[2])) % P,
            (a[0] * b[2] + a[1] * b[1] + a[2] * b[0] + nb * (a[3] * b[3])) % P,
            (a[0] * b[3] + a[1] * b[2] + a[2] * b[1] + a[3] * b[0]) % P)


def ext_add(a, b):
    return tuple((x + y) % P for x, y in zip(a, b))


class Checksum:
    """prove/witgen/poseidon2.rs:236-285"""

    def __init__(self, rand_z):
        cur = (1, 0, 0, 0)
        self.powers = []
        for _ in range(17):
            self.powers.append(cur)
            cur = ext_mul(cur, rand_z)
        self.zcheck = (0, 0, 0, 0)

    def start(self):
        self.zcheck = ext_mul(self.zcheck, self.powers[16])

    def clear(self):
        self.zcheck = (0, 0, 0, 0)

    def add(self, tx_kind, idx, txn):
        addr, cycle, word, prev_cycle, prev_word = txn
        if tx_kind == TX_READ:
            c0, c1 = 0, 1
        elif tx_kind == TX_PAGE_IN:
            c0, c1 = 0, cycle - prev_cycle
        else:
            c0, c1 = (word & 0xffff) - (prev_word & 0xffff), (word >> 16) - (prev_word >> 16)
        # the reference computes these in i32 and adds P once when negative
        c0 = _i32(c0)
        c1 = _i32(c1)
        if c0 < 0:
            c0 += P
        if c1 < 0:
            c1 += P
        c0 %= P
        c1 %= P
        self.zcheck = ext_add(self.zcheck, ext_mul(self.powers[2 * idx], (c0, 0, 0, 0)))
        self.zcheck = ext_add(self.zcheck, ext_mul(self.powers[2 * idx + 1], (c1, 0, 0, 0)))


def _i32(x):
    x &= M32
    return x - (1 << 32) if x & 0x80000000 else x


# ---- Segment + Executor::run (execute/executor.rs) -------------------------------------------------------------------
class Segment:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class _ExecCtx:
    """Risc0Context for Executor (executor.rs:507-660)"""

    def __init__(self, image, input_digest, read_fn, write_fn):
        self.pager = PagedMemory(image)
        self.pc = 0
        self.machine_mode = 0
        self.user_cycles = 0
        self.terminate_state = None
        self.output_digest = None
        self.read_record, self.write_record = [], []
        self.input_digest = tuple(input_digest)
        self.read_fn, self.write_fn = read_fn, write_fn

    def load_u32(self, op, waddr):
        return self.pager.peek(waddr) if op == PEEK else self.pager.load(waddr)

    def store_u32(self, waddr, word):
        self.pager.store(waddr, word)

    def resume_hook(self):
        for i, w in enumerate(self.input_digest):
            self.store_u32(GLOBAL_INPUT_ADDR // 4 + i, w)

    def suspend_hook(self):
        pass

    def trap_rewind(self):
        pass

    def on_insn_end(self, kind):
        self.user_cycles += 1

    def on_ecall_cycle(self, cur, nxt, s0, s1, s2):
        self.user_cycles += 1

    def on_poseidon2_cycle(self, cur, p2):
        self.user_cycles += 1

    def on_sha2_cycle(self, cur, sha2):
        self.user_cycles += 1

    def ecall_bigint(self):
        # executor.rs:652-656, execute/bigint.rs:150-170: run the nondet program, put its results into guest memory
        _mode, _ptr, verify_size, witness = bigint_ecall(self)
        for addr in sorted(witness):
            for i in range(BIGINT_WIDTH_WORDS):
                self.store_u32(addr + i, int.from_bytes(witness[addr][4 * i:4 * i + 4], "little"))
        self.user_cycles += verify_size + 1

    def on_terminate(self, a0, a1):
        self.terminate_state = (a0, a1)
        self.output_digest = tuple(self.load_u32(PEEK, GLOBAL_OUTPUT_ADDR // 4 + i) for i in range(8))

    def host_read(self, fd, n):
        data = bytes(self.read_fn(fd, n))[:n]
        self.read_record.append(data)
        return data

    def host_write(self, fd, data):
        rlen = self.write_fn(fd, data)
        self.write_record.append(rlen)
        return rlen


def execute(image, segment_po2=20, max_insn_cycles=None, max_cycles=1 << 24, input_digest=(0,) * 8, read_fn=None,
            write_fn=None, max_segments=None, povw_nonce=None):
    """Executor::run (executor.rs:209-405): runs the guest, returns the list of Segments. NullSyscall-style defaults."""
    if max_insn_cycles is None:
        max_insn_cycles = MAX_INSN_CYCLES if segment_po2 >= 15 else MAX_INSN_CYCLES_LOWER_PO2
    read_fn = read_fn or (lambda fd, n: bytes(i & 0xff for i in range(n)))
    write_fn = write_fn or (lambda fd, data: len(data))
    segment_limit = 1 << segment_po2
    assert max_insn_cycles < segment_limit
    segment_threshold = segment_limit - max_insn_cycles
    ctx = _ExecCtx(image, input_digest, read_fn, write_fn)
    m = Machine(ctx)
    existing = image.clone()
    segments = []
    total_user = 0

    def seg_cycles():
        return ctx.user_cycles + ctx.pager.cycles + RESERVED_CYCLES

    def emit(po2, threshold):
        partial_pages = ctx.pager.commit()
        page_indexes = ctx.pager.page_indexes()
        pre = existing.image_id()
        partial = compute_partial_image(existing, page_indexes)
        for idx, page in partial_pages.items():
            existing.set_page(idx, page)
        existing.update_digests()
        segments.append(Segment(partial_image=partial, pre_state=pre, post_state=existing.image_id(),
                                input=ctx.input_digest, output=ctx.output_digest, terminate_state=ctx.terminate_state,
                                read_record=ctx.read_record, write_record=ctx.write_record, suspend_cycle=ctx.user_cycles,
                                paging_cycles=ctx.pager.cycles, po2=po2, index=len(segments),
                                segment_threshold=threshold,
                                povw_nonce=tuple(povw_nonce) if povw_nonce is not None else None))
        ctx.read_record, ctx.write_record = [], []

    m.resume()
    while ctx.terminate_state is None:
        if total_user + ctx.user_cycles >= max_cycles:
            raise RuntimeError("Session limit exceeded")
        if seg_cycles() > segment_threshold:
            assert seg_cycles() < segment_limit, "segment limit too small for instruction at pc %#x" % ctx.pc
            m.suspend()
            emit(segment_po2, segment_threshold)
            if max_segments is not None and len(segments) >= max_segments:
                return segments
            total_user += ctx.user_cycles
            ctx.user_cycles = 0
            ctx.pager.reset()
            m.resume()
        m.step()
    m.suspend()
    final_cycles = 1 << max(seg_cycles() - 1, 0).bit_length()
    emit(final_cycles.bit_length() - 1, 0)
    return segments


# ---- Preflight (prove/witgen/preflight.rs) ---------------------------------------------------------------------------
CYCLE_DTYPE = np.dtype([("state", "<u4"), ("pc", "<u4"), ("major", "u1"), ("minor", "u1"), ("machine_mode", "u1"),
                        ("padding", "u1"), ("user_cycle", "<u4"), ("txn_idx", "<u4"), ("paging_idx", "<u4"),
                        ("bigint_idx", "<u4"), ("diff_count", "<u4", (2,))])
TXN_DTYPE = np.dtype([("addr", "<u4"), ("cycle", "<u4"), ("word", "<u4"), ("prev_cycle", "<u4"), ("prev_word", "<u4")])
assert CYCLE_DTYPE.itemsize == 36 and TXN_DTYPE.itemsize == 20


class _Preflight:
    def __init__(self, segment, rand_z, write_record_off_by_one=True):
        self.segment = segment
        self.write_record_off_by_one = write_record_off_by_one
        self.rand_z = tuple(int(x) for x in rand_z)
        self.cycles = []     # [state, pc, major, minor, machine_mode, user_cycle, txn_idx, paging_idx, bigint_idx, d0, d1]
        self.backs = []      # None | ("ecall", s0, s1, s2) | ("p2", Poseidon2State) | ("sha2", Sha2State) | ("bigint", BigIntState)
        self.txns = []       # [addr, cycle, word, prev_cycle, prev_word]
        self.bigint_bytes = bytearray()   # 16 bytes per bigint cycle (preflight.rs:402-404), indexed by the cycle's bigint_idx
        self.bigint_idx = 0
        self.pager = PagedMemory(segment.partial_image)
        self.pc = 0
        self.machine_mode = 0
        self.user_cycle = 0
        self.user_cycles = 0
        self.txn_idx = 0
        self.cur_read = self.cur_write = 0
        self.orig_words = {}
        self.prev_cycle = {}
        self.page_memory = {}
        for node_idx, digest in segment.partial_image.digests.items():
            base = node_idx_to_addr(node_idx)
            for i in range(8):
                self.page_memory[base + i] = digest[i]

    # -- cycles
    def add_cycle(self, state, pc, major, minor, paging_idx, back):
        self.cycles.append([state, pc, major, minor, self.machine_mode, self.user_cycle, self.txn_idx, paging_idx, self.bigint_idx,
                            0, 0])
        self.backs.append(back)
        self.txn_idx = len(self.txns)
        self.bigint_idx = len(self.bigint_bytes)

    def add_cycle_special(self, cur_state, next_state, pc, paging_idx, back):
        self.add_cycle(next_state, pc, 7 + cur_state // 8, cur_state % 8, paging_idx, back)

    def on_insn_end(self, kind):
        if kind == EANY:
            if self.cycles[-1][4] != 0:
                self.add_cycle(CS.Decode, self.pc, MAJOR_ECALL0, 0, 0, None)
            else:
                self.add_cycle(CS.Decode, self.pc, MAJOR_CONTROL0, CONTROL_USER_ECALL, 0, None)
        elif kind == MRET:
            self.add_cycle(CS.Decode, self.pc, MAJOR_CONTROL0, CONTROL_MRET, 0, None)
        elif kind == FENCE:
            self.add_cycle(CS.Decode, self.pc, MAJOR_CONTROL0, CONTROL_FENCE, 0, None)
        else:
            self.add_cycle(CS.Decode, self.pc, kind // 8, kind % 8, 0, None)
        self.user_cycle += 1
        self.user_cycles += 1

    def on_ecall_cycle(self, cur, nxt, s0, s1, s2):
        self.add_cycle_special(cur, nxt, self.pc, 0, ("ecall", s0, s1, s2))
        self.user_cycles += 1

    def on_poseidon2_cycle(self, cur_state, p2):
        self.add_cycle_special(cur_state, p2.next_state, self.pc, node_addr_to_idx(p2.buf_out_addr), ("p2", p2.clone()))
        self.user_cycles += 1

    def ecall_bigint(self):                        # preflight.rs:699-701
        bigint_ecall_preflight(self)

    def on_bigint_cycle(self, cur_state, st):      # preflight.rs:471-481
        self.bigint_bytes += st.bytes
        self.add_cycle_special(cur_state, st.next_state, self.pc, 0, ("bigint", st.clone()))
        self.user_cycles += 1

    def on_sha2_cycle(self, cur_state, sha2):      # preflight.rs:677-686
        self.add_cycle_special(cur_state, sha2.next_state, self.pc, node_addr_to_idx(sha2.state_out_addr), ("sha2", sha2.clone()))
        self.user_cycles += 1

    def on_terminate(self, a0, a1):
        pass

    def trap_rewind(self):
        del self.txns[self.txn_idx:]

    def host_read(self, fd, n):
        rec = self.segment.read_record[self.cur_read]
        assert len(rec) <= n, "Invalid segment: truncated read record"
        self.cur_read += 1
        return rec

    def host_write(self, fd, data):
        # preflight.rs:666-674 increments cur_write BEFORE indexing the record (sic), so the reference's own preflight
        # cannot get through a segment whose last ecall is a host write. write_record_off_by_one=False reads the entry
        # the executor recorded for THIS write instead, which is what lets the HostWrite arm of the circuit be exercised.
        self.cur_write += 1
        if self.write_record_off_by_one:
            return self.segment.write_record[self.cur_write]
        return self.segment.write_record[self.cur_write - 1]

    # -- memory
    def load_u32(self, op, waddr):
        if op == PEEK:
            return self.pager.peek(waddr)
        cycle = 2 * len(self.cycles)
        if waddr >= MERKLE_TREE_START_WADDR:
            if waddr < MERKLE_TREE_END_WADDR:
                if waddr not in self.page_memory:
                    raise KeyError("Invalid load from page memory %#x" % waddr)
                word = self.page_memory[waddr]
            elif POVW_NONCE_START_WADDR <= waddr < POVW_NONCE_END_WADDR:
                nonce = self.segment.povw_nonce or (0,) * 8
                word = nonce[waddr - POVW_NONCE_START_WADDR]
            else:
                raise ValueError("invalid memory access in special region: %#x" % waddr)
        else:
            word = self.pager.load(waddr)
        if op == RECORD:
            self.orig_words.setdefault(waddr, word)
            prev = self.prev_cycle.get(waddr, M32)
            self.prev_cycle[waddr] = cycle
            self.txns.append([waddr, cycle, word, prev, word])
        return word

    def store_u32(self, waddr, word):
        cycle = 2 * len(self.cycles) + 1
        if waddr >= MEMORY_END_WADDR:
            if waddr not in self.page_memory:
                raise KeyError("Invalid store to page memory %#x" % waddr)
            prev_word = self.page_memory[waddr]
            self.page_memory[waddr] = word
        else:
            prev_word = self.pager.load(waddr)
            self.pager.store(waddr, word)
        prev = self.prev_cycle.get(waddr, M32)
        self.prev_cycle[waddr] = cycle
        self.txns.append([waddr, cycle, word, prev, prev_word])

    # -- Risc0Context hooks
    def resume_hook(self):
        self.add_cycle_special(CS.Resume, CS.Resume, self.pc, 0, None)
        for i, w in enumerate(self.segment.input):
            self.store_u32(GLOBAL_INPUT_ADDR // 4 + i, w)
        self.add_cycle_special(CS.Resume, CS.Decode, self.pc, 0, None)

    def suspend_hook(self):
        self.pc = 0
        self.add_cycle_special(CS.Suspend, CS.Suspend, 0, 0, None)
        for i in range(8):
            self.load_u32(RECORD, GLOBAL_OUTPUT_ADDR // 4 + i)
        self.machine_mode = 3
        self.add_cycle_special(CS.Suspend, CS.PoseidonEntry, 0, 0, None)

    # -- the seven passes of Segment::preflight (preflight.rs:92-113)
    def run(self):
        seg = self.segment
        # read_povw_nonce
        for i in range(8):
            self.load_u32(RECORD, POVW_NONCE_START_WADDR + i)
        self.add_cycle_special(CS.LoadRootAndNonce, CS.LoadRootAndNonce, 0, 0, None)
        # read_pages
        for i in range(8):
            self.load_u32(RECORD, get_digest_addr(1) + i)
        self.add_cycle_special(CS.LoadRootAndNonce, CS.PoseidonEntry, 0, 0, None)
        activity = self.pager.loaded_pages()
        self.on_poseidon2_cycle(CS.PoseidonEntry, Poseidon2State.new_start(0))
        for node_idx in activity.nodes:
            Poseidon2State.new_node(node_idx, True).rest(self, CS.PoseidonPaging)
        self.machine_mode = 1
        for page_idx in activity.pages:
            Poseidon2State.new_page(page_idx, True).rest(self, CS.PoseidonPaging)
        self.machine_mode = 2
        self.on_poseidon2_cycle(CS.PoseidonPaging, Poseidon2State.new_done(MERKLE_TREE_START_WADDR, CS.Resume, 2))
        self.user_cycles = 0
        # body
        m = Machine(self)
        m.resume()
        while self.user_cycles < seg.suspend_cycle:
            m.step()
        m.suspend()
        # write_pages
        activity = self.pager.dirty_pages()
        self.pager.commit()
        self.on_poseidon2_cycle(CS.PoseidonEntry, Poseidon2State.new_start(3))
        for page_idx in reversed(activity.pages):
            Poseidon2State.new_page(page_idx, False).rest(self, CS.PoseidonPaging)
        self.machine_mode = 4
        for node_idx in reversed(activity.nodes):
            Poseidon2State.new_node(node_idx, False).rest(self, CS.PoseidonPaging)
        self.machine_mode = 5
        self.on_poseidon2_cycle(CS.PoseidonPaging, Poseidon2State.new_done(MERKLE_TREE_END_WADDR, CS.StoreRoot, 5))
        self.machine_mode = 0
        for i in range(8):
            self.load_u32(RECORD, get_digest_addr(1) + i)
        self.add_cycle_special(CS.StoreRoot, CS.ControlTable, 0, 0, None)
        # generate_tables / fini
        table_split_cycle = len(self.cycles)
        start = len(self.cycles)
        for i in range(16, 256, 16):
            self.add_cycle_special(CS.ControlTable, CS.ControlTable, i, 0, None)
        self.machine_mode = 1
        for i in range(0, 64 * 1024, 16):
            self.add_cycle_special(CS.ControlTable, CS.ControlTable, i, 0, None)
        self.machine_mode = 0
        self.add_cycle_special(CS.ControlTable, CS.ControlDone, 0, 0, None)
        if seg.terminate_state is None:
            if len(self.cycles) < seg.segment_threshold:
                raise RuntimeError("Stopping segment too early")
            diff = len(self.cycles) - seg.segment_threshold
            self.cycles[diff // 2][9 + diff % 2] += 1
        self.machine_mode = 1
        self.add_cycle_special(CS.ControlDone, CS.ControlDone, 0, 0, None)
        assert len(self.cycles) - start == RESERVED_CYCLES
        last_cycle = 1 << seg.po2
        assert len(self.cycles) <= last_cycle, "cycles <= 1 << segment.po2"
        while len(self.cycles) < last_cycle:
            self.add_cycle_special(CS.ControlDone, CS.ControlDone, 0, 0, None)
        # wrap_memory_txns
        for txn in self.txns:
            addr = txn[0]
            if txn[3] == M32:
                txn[3] = self.prev_cycle[addr]
            else:
                assert txn[1] != txn[3]
                diff = txn[1] - 1 - txn[3]
                self.cycles[diff // 2][9 + diff % 2] += 1
            if txn[1] == self.prev_cycle[addr]:
                txn[2] = self.orig_words.get(addr, 0)
        # update_p2_zcheck
        checksum = Checksum(self.rand_z)
        for row, back in enumerate(self.backs):
            if back is not None and back[0] == "p2":
                p2 = back[1]
                cyc = self.cycles[row]
                state = (cyc[2] - 7) * 8 + cyc[3]
                if state == CS.PoseidonLoadIn:
                    checksum.start()
                    for i, t in enumerate(range(cyc[6], self.cycles[row + 1][6])):
                        checksum.add(p2.load_tx_type, i, self.txns[t])
                if state in (CS.PoseidonLoadIn, CS.PoseidonExtRound, CS.PoseidonIntRound):
                    p2.zcheck = checksum.zcheck
                else:
                    checksum.clear()
        return table_split_cycle


# ---- layout lookups (from the committed circuit IR) ------------------------------------------------------------------
_IR = None


def circuit_ir():
    global _IR
    if _IR is None:
        _IR = json.load(gzip.open(os.path.join(HERE, "circuits", "rv32im_witgen.ir.json.gz")))
    return _IR


def _flat_size(ty, types, cache={}):
    if ty == "Reg":
        return 1
    if ty not in cache:
        k = types[ty]
        cache[ty] = sum(_flat_size(f[1], types) for f in k[1]) if k[0] == "struct" else k[2] * _flat_size(k[1], types)
    return cache[ty]


def layout_col(layout, path):
    """column of a register: layout_col("kLayout_Top", "instResult.arm8.s0._super")"""
    ir = circuit_ir()
    types = ir["types"]
    lay = ir["layouts"][layout]
    ty, off = lay["type"], 0
    import re
    for step in re.findall(r"[A-Za-z_][A-Za-z_0-9]*|\[\d+\]", path):
        k = types[ty]
        if step.startswith("["):
            assert k[0] == "array"
            off += int(step[1:-1]) * _flat_size(k[1], types)
            ty = k[1]
        else:
            assert k[0] == "struct", (ty, step)
            for fname, fty in k[1]:
                if fname == step:
                    ty = fty
                    break
                off += _flat_size(fty, types)
            else:
                raise KeyError("%s has no field %s" % (ty, step))
    assert ty == "Reg", "path does not end at a register: %s" % path
    return lay["cols"][off]


def montgomery(x):
    """u32 array of normal-form integers -> Montgomery words"""
    return (np.asarray(x, dtype=np.uint64) % P * ((1 << 32) % P) % P).astype(np.uint32)


class Injector:
    """witgen/mod.rs:320-370: CSR scatter (index per row, word offsets col * rows + row, Montgomery values)"""

    def __init__(self, rows):
        self.rows = rows
        self.offsets, self.values, self.index = [], [], [0]

    def set(self, row, col, value):
        self.offsets.append(col * self.rows + row)
        self.values.append(value)

    def set_u32_bits(self, row, col, value):
        for i in range(32):
            self.set(row, col + i, (value >> i) & 1)

    def push(self):
        self.index.append(len(self.offsets))

    def arrays(self):
        return (np.asarray(self.index, dtype=np.uint32), np.asarray(self.offsets, dtype=np.uint32),
                montgomery(np.asarray(self.values, dtype=np.uint64)))


class PreflightResults:
    """witgen/mod.rs:55-88: what prove_core starts from. cycles / txns are the RawPreflightTrace arrays."""

    def __init__(self, segment, rand_z, write_record_off_by_one=True):
        pf = _Preflight(segment, rand_z, write_record_off_by_one)
        self.table_split_cycle = pf.run()
        self.po2 = segment.po2
        self.rows = 1 << segment.po2
        self.segment = segment
        self.rand_z = pf.rand_z
        n = len(pf.cycles)
        assert n == self.rows
        cyc = np.zeros(n, dtype=CYCLE_DTYPE)
        arr = np.asarray(pf.cycles, dtype=np.uint64)
        for i, f in enumerate(("state", "pc", "major", "minor", "machine_mode", "user_cycle", "txn_idx", "paging_idx", "bigint_idx")):
            cyc[f] = arr[:, i]
        cyc["diff_count"][:, 0] = arr[:, 9]
        cyc["diff_count"][:, 1] = arr[:, 10]
        self.cycles = cyc
        tx = np.zeros(len(pf.txns), dtype=TXN_DTYPE)
        tarr = np.asarray(pf.txns, dtype=np.uint64).reshape(-1, 5)
        for i, f in enumerate(("addr", "cycle", "word", "prev_cycle", "prev_word")):
            tx[f] = tarr[:, i]
        self.txns = tx
        self.bigint_bytes = np.frombuffer(bytes(pf.bigint_bytes), dtype=np.uint8).copy()
        self.backs = pf.backs
        self.has_bigint = any(b is not None and b[0] == "bigint" for b in pf.backs)
        self.user_cycles = segment.suspend_cycle
        self.injector = self._build_injector(pf)
        self.global_ = self._build_global()

    def _build_injector(self, pf):
        T = "kLayout_Top"
        ecall = [layout_col(T, "instResult.arm8.%s._super" % s) for s in ("s0", "s1", "s2")]
        st = "instResult.arm9.state."
        names = ("hasState", "stateAddr", "bufOutAddr", "isElem", "checkOut", "loadTxType", "nextState", "subState", "bufInAddr",
                 "count", "mode")
        p2_cols = [layout_col(T, st + nm + "._super") for nm in names]
        p2_cols += [layout_col(T, st + "inner[%d]._super" % i) for i in range(24)]
        z = layout_col(T, st + "zcheck._super")
        p2_cols += [z, z + 1, z + 2, z + 3]
        sst = "instResult.arm11.state."
        sha_fp = [layout_col(T, sst + nm + "._super") for nm in ("stateInAddr", "stateOutAddr", "dataAddr", "count", "kAddr", "round",
                                                                  "nextState")]
        sha_bits = [layout_col(T, sst + "%s[0]._super" % nm) for nm in ("a", "e", "w")]
        bst = "instResult.arm12.state."
        big_cols = [layout_col(T, bst + nm + "._super") for nm in ("isEcall", "mode", "pc", "polyOp", "coeff")]
        big_cols += [layout_col(T, bst + "bytes[%d]._super" % i) for i in range(16)] + [layout_col(T, bst + "nextState._super")]
        cycle_col = layout_col(T, "cycle._super")
        pc_low, pc_high = layout_col(T, "nextPcLow._super"), layout_col(T, "nextPcHigh._super")
        next_state, next_mm = layout_col(T, "nextState_0._super"), layout_col(T, "nextMachineMode._super")
        inj = Injector(self.rows)
        for row, back in enumerate(pf.backs):
            cyc = pf.cycles[row]
            if back is not None:
                if back[0] == "ecall":
                    for col, v in zip(ecall, back[1:4]):
                        inj.set(row, col, v)
                elif back[0] == "bigint":         # witgen/mod.rs:261-265
                    for col, v in zip(big_cols, back[1].as_array()):
                        inj.set(row, col, v)
                elif back[0] == "sha2":           # witgen/mod.rs:253-260: 7 field columns, then a / e / w bit by bit
                    for col, v in zip(sha_fp, back[1].fp_array()):
                        inj.set(row, col, v)
                    for col, v in zip(sha_bits, back[1].u32_array()):
                        for i in range(32):
                            inj.set(row, col + i, (v >> i) & 1)
                else:
                    for col, v in zip(p2_cols, back[1].as_array()):
                        inj.set(row, col, v)
            inj.set(row, cycle_col, row)
            inj.set(row, pc_low, cyc[1] & 0xffff)
            inj.set(row, pc_high, cyc[1] >> 16)
            inj.set(row, next_state, cyc[0])
            inj.set(row, next_mm, cyc[4])
            inj.push()
        return inj.arrays()

    def bigint_accum_injector(self, mix):
        """witgen/mod.rs:186-207: the BigIntAccumState cells of every bigint cycle for the transcript's accum mix (36
        Montgomery words; the last four are the evaluation point) -> CSR scatter arrays for the ACCUM matrix, or None
        when the segment has no bigint cycles"""
        if not self.has_bigint:
            return None
        rinv = pow(1 << 32, -1, P)
        point = tuple(int(w) * rinv % P for w in np.asarray(mix, dtype=np.uint64)[-4:])
        A = "kLayout_TopAccum"
        cols = [layout_col(A, "user._0.state.%s._super" % nm) + i for nm in ("poly", "term", "total") for i in range(4)]
        acc, inj = BigIntAccum(point), Injector(self.rows)
        for row, back in enumerate(self.backs):
            if back is not None and back[0] == "bigint":
                acc.step(back[1])
                for col, v in zip(cols, acc.as_array()):
                    inj.set(row, col, v)
                inj.push()
        return inj.arrays()

    def _build_global(self):
        """witgen/mod.rs:272-318; INVALID where the witness generator fills the value"""
        G = "kLayoutGlobal"
        seg = self.segment
        g = np.full(circuit_ir()["regcounts"]["kRegCountGlobal"], 0xFFFFFFFF, dtype=np.uint32)

        def put(col, v):
            g[col] = montgomery([v])[0]

        for i, w in enumerate(seg.pre_state):
            put(layout_col(G, "stateIn.values[%d].low._super" % i), w & 0xffff)
            put(layout_col(G, "stateIn.values[%d].high._super" % i), w >> 16)
        for i, w in enumerate(seg.input):
            put(layout_col(G, "input.values[%d].low._super" % i), w & 0xffff)
            put(layout_col(G, "input.values[%d].high._super" % i), w >> 16)
        rng = layout_col(G, "rng._super")
        for i, e in enumerate(self.rand_z):
            put(rng + i, e)
        put(layout_col(G, "isTerminate._super"), int(seg.terminate_state is not None))
        put(layout_col(G, "shutdownCycle._super"), seg.segment_threshold)
        nonce = seg.povw_nonce or (0,) * 8
        for i, w in enumerate(nonce):
            put(layout_col(G, "povwNonce.values[%d].low._super" % i), w & 0xffff)
            put(layout_col(G, "povwNonce.values[%d].high._super" % i), w >> 16)
        return g


# ---- a tiny assembler for hand-written guests (execute/testutil.rs:186-348) ------------------------------------------
class Assembler:
    def __init__(self, base=USER_START_ADDR + WORD_SIZE):
        self.text, self.data, self.base = [], {}, base

    def program(self):
        entry = self.base
        image = {entry + 4 * i: w for i, w in enumerate(self.text)}
        image.update(self.data)
        return entry, image

    def word(self, addr, word):
        self.data[addr] = word

    def _i(self, imm, rs1, f3, rd, op):
        self.text.append(((imm << 20) | (rs1 << 15) | (f3 << 12) | (rd << 7) | op) & M32)

    def _r(self, f7, rs2, rs1, f3, rd, op):
        self.text.append((f7 << 25) | (rs2 << 20) | (rs1 << 15) | (f3 << 12) | (rd << 7) | op)

    def _b(self, imm, rs2, rs1, f3):
        imm &= M32
        self.text.append((((((imm >> 12) & 1) << 6) | ((imm >> 5) & 0x3f)) << 25) | (rs2 << 20) | (rs1 << 15) | (f3 << 12) |
                         (((((imm >> 1) & 0xf) << 1) | ((imm >> 11) & 1)) << 7) | 0b1100011)

    def _s(self, imm, rs2, rs1, f3):
        imm &= 0xfff
        self.text.append(((imm >> 5) << 25) | (rs2 << 20) | (rs1 << 15) | (f3 << 12) | ((imm & 31) << 7) | 0b0100011)

    def addi(self, rd, rs1, imm):
        self._i(imm & 0xfff, rs1, 0, rd, 0b0010011)

    def add(self, rd, rs1, rs2):
        self._r(0, rs2, rs1, 0, rd, 0b0110011)

    def op(self, f7, f3, rd, rs1, rs2):
        self._r(f7, rs2, rs1, f3, rd, 0b0110011)

    def opi(self, f3, rd, rs1, imm):
        self._i(imm & 0xfff, rs1, f3, rd, 0b0010011)

    def lui(self, rd, imm20):
        self.text.append(((imm20 << 12) | (rd << 7) | 0b0110111) & M32)

    def li(self, rd, imm):
        if imm < (1 << 11):
            self.addi(rd, 0, imm)
        else:
            low = ((imm & 0xfff) ^ 0x800) - 0x800
            high = ((imm - low) >> 12) & 0xfffff
            self.lui(rd, high)
            self.addi(rd, rd, low)

    def blt(self, rs1, rs2, off):
        self._b(off, rs2, rs1, 4)

    def beq(self, rs1, rs2, off):
        self._b(off, rs2, rs1, 0)

    def bne(self, rs1, rs2, off):
        self._b(off, rs2, rs1, 1)

    def lw(self, rd, rs1, imm):
        self._i(imm & 0xfff, rs1, 2, rd, 0b0000011)

    def lb(self, rd, rs1, imm):
        self._i(imm & 0xfff, rs1, 0, rd, 0b0000011)

    def load(self, f3, rd, rs1, imm):
        self._i(imm & 0xfff, rs1, f3, rd, 0b0000011)

    def sw(self, rs2, rs1, imm):
        self._s(imm, rs2, rs1, 2)

    def store(self, f3, rs2, rs1, imm):
        self._s(imm, rs2, rs1, f3)

    def ecall(self):
        self._i(0, 0, 0, 0, 0b1110011)

    def mret(self):
        self.text.append(0x30200073)

    def host_terminate(self, a0, a1):
        self.li(REG_A7, HOST_ECALL_TERMINATE)
        self.li(REG_A0, a0)
        self.li(REG_A1, a1)
        self.ecall()

    def host_ecall_read(self, fd, ptr, ln):
        self.li(REG_A7, HOST_ECALL_READ)
        self.li(REG_A0, fd)
        self.li(REG_A1, ptr)
        self.li(REG_A2, ln)
        self.ecall()


def simple_loop_kernel(count):
    """execute/testutil.rs:152-161 kernel::simple_loop: the reference's loop guest (2 instructions per iteration)"""
    a4, a5 = 14, 15
    asm = Assembler()
    asm.addi(a4, 0, 0)
    asm.li(a5, count)
    asm.addi(a4, a4, 1)
    asm.blt(a4, a5, -4)
    asm.host_terminate(0, 0)
    entry, image = asm.program()
    return MemoryImage.new_kernel(entry, image)


def user_mode_guest(count=20):
    """a user-mode loop under a minimal machine-mode kernel: the kernel's entry points MEPC at the user program and
    `mret`s into it; the user program counts and issues `ecall`, which traps to the kernel's dispatch address
    (ECALL_DISPATCH_ADDR, r0vm.rs:372-381), whose handler terminates. Exercises the user register file, USER_ECALL and
    MRET control cycles and the machine-mode switches."""
    a4, a5, t0, t1 = 14, 15, 5, 6
    user = Assembler()
    user.addi(a4, 0, 0)
    user.li(a5, count)
    user.addi(a4, a4, 1)
    user.blt(a4, a5, -4)
    user.ecall()
    user_entry, user_image = user.program()
    kern = Assembler(base=KERNEL_START_ADDR)
    kern.li(t1, user_entry - 4)          # mret resumes at MEPC + 4
    kern.li(t0, MEPC_ADDR)
    kern.sw(t1, t0, 0)
    kern.mret()
    handler = KERNEL_START_ADDR + 0x100
    while KERNEL_START_ADDR + 4 * len(kern.text) < handler:
        kern.text.append(0x00000013)     # nop
    kern.host_terminate(0, 0)
    kentry, kimage = kern.program()
    image = dict(user_image)
    image.update(kimage)
    image[ECALL_DISPATCH_ADDR] = handler
    return MemoryImage.new_kernel(kentry, image)


SHA256_K = [
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2]
SHA256_IV = [0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19]
SHA2_GUEST_OUT_ADDR = 0x00500400


def sha2_guest(message=b"abc", loops=3, repeat=1):
    """machine-mode guest that hashes `message` with the sha2 ecall (r0vm.rs:559-571, execute/sha2.rs): the padded
    message's blocks, the SHA-256 initial state (big-endian words, as the ecall reads them) and the round-constant table
    live in guest memory; the digest lands at SHA2_GUEST_OUT_ADDR as 32 bytes in digest order. A short loop before and a
    load of the result after the ecall put ordinary cycles on both sides of the SHA cycles."""
    a4, a5, t3 = 14, 15, 28
    data_addr, state_addr, k_addr = 0x00500000, 0x00500300, 0x00500500
    ml = len(message)
    padded = bytes(message) + b"\x80" + b"\x00" * ((55 - ml) % 64) + (8 * ml).to_bytes(8, "big")
    assert len(padded) % 64 == 0 and len(padded) // 64 <= MAX_SHA_COUNT
    asm = Assembler()
    for i in range(len(padded) // 4):
        asm.word(data_addr + 4 * i, int.from_bytes(padded[4 * i:4 * i + 4], "little"))
    for i, w in enumerate(SHA256_IV):
        asm.word(state_addr + 4 * i, _bswap(w))
    for i, w in enumerate(SHA256_K):
        asm.word(k_addr + 4 * i, w)
    asm.addi(a4, 0, 0)
    asm.li(a5, loops)
    asm.addi(a4, a4, 1)
    asm.blt(a4, a5, -4)
    for _ in range(repeat):              # repeat > 1: the same hash again and again (long sessions that split)
        asm.li(REG_A0, state_addr)
        asm.li(REG_A1, SHA2_GUEST_OUT_ADDR)
        asm.li(REG_A2, data_addr)
        asm.li(REG_A3, len(padded) // 64)
        asm.li(REG_A4, k_addr)
        asm.li(REG_A7, HOST_ECALL_SHA2)
        asm.ecall()
    asm.li(t3, SHA2_GUEST_OUT_ADDR)
    asm.load(2, a4, t3, 0)               # lw: the stored state is read back by an ordinary instruction
    asm.load(2, a5, t3, 28)
    asm.host_terminate(0, 0)
    entry, image = asm.program()
    return MemoryImage.new_kernel(entry, image)


BIGINT_GUEST_OUT_ADDR = 0x005010c0


def bigint_guest(blob, inputs, outputs, loops=3):
    """machine-mode guest that runs a bigint2 blob (header: nondet / verify / consts / temp sizes in words, then the bibc
    program, the verify program and the constants; zkvm/platform/src/syscall.rs:1044-1115 shows how a guest lays the
    registers out). inputs: {register: bytes} operand regions (a program addresses them as register + 16-byte offset),
    outputs: {register: size in bytes} result regions; the scratch space the blob asks for hangs off sp.
    Returns (image, {register: byte address})."""
    import struct
    nondet_size, verify_size, consts_size, temp_size = struct.unpack_from("<4I", blob, 0)
    assert len(blob) == 16 + 4 * (nondet_size + verify_size + consts_size)
    blob_addr, next_addr = 0x00500000, 0x00501000
    asm = Assembler()
    for i in range(len(blob) // 4):
        asm.word(blob_addr + 4 * i, int.from_bytes(blob[4 * i:4 * i + 4], "little"))
    where = {}
    for reg, data in inputs.items():
        assert len(data) % 16 == 0
        where[reg] = next_addr
        for i in range(len(data) // 4):
            asm.word(next_addr + 4 * i, int.from_bytes(data[4 * i:4 * i + 4], "little"))
        next_addr += len(data) + 32
    for reg, size in list(outputs.items()) + [(REG_SP, max(16, 4 * temp_size))]:
        where[reg] = next_addr
        next_addr += -(-size // 16) * 16 + 32
    a4, a5, t4 = 14, 15, 29
    asm.addi(a4, 0, 0)
    asm.li(a5, loops)
    asm.addi(a4, a4, 1)
    asm.blt(a4, a5, -4)
    asm.li(REG_T0, 0)                                   # mode
    asm.li(REG_A0, blob_addr)
    asm.li(REG_T1, blob_addr + 16)
    asm.li(REG_T2, blob_addr + 16 + 4 * nondet_size)
    asm.li(REG_T3, blob_addr + 16 + 4 * (nondet_size + verify_size))
    for reg, addr in where.items():
        asm.li(reg, addr)
    asm.li(REG_A7, HOST_ECALL_BIGINT)
    asm.ecall()
    first_out = where[next(iter(outputs))]
    asm.li(t4, first_out)
    asm.load(2, a5, t4, 0)                              # the result is read back by an ordinary instruction
    asm.host_terminate(0, 0)
    entry, image = asm.program()
    return MemoryImage.new_kernel(entry, image), where


def bigint_modmul_guest(blob, a, b, n, loops=3):
    """`modmul_256`: a, b, n are 256-bit operands at a1, a2, a3, the product a * b mod n lands at a4"""
    image, where = bigint_guest(blob, {REG_A1: a.to_bytes(32, "little"), REG_A2: b.to_bytes(32, "little"),
                                       REG_A3: n.to_bytes(32, "little")}, {REG_A4: 32}, loops)
    assert where[REG_A4] == BIGINT_GUEST_OUT_ADDR
    return image


P2_GUEST_OUT_ADDR, P2_GUEST_STATE_ADDR = 0x00502000, 0x00502100


def poseidon2_ecall_guest(words, is_elem, state=None, expect=None):
    """machine-mode guest that absorbs `words` with the poseidon2 ecall (r0vm.rs:545-557, execute/poseidon2.rs:56-72,
    285-293; registers carry byte addresses, see Machine._machine_ecall): is_elem - 16 field elements per block, else 8 words = 16 half-words per
    block; state - 8 capacity words loaded before / stored after (a0 != 0); expect - a digest already in memory that
    the ecall checks instead of storing (PFLAG_CHECK_OUT)."""
    per_block = 16 if is_elem else 8
    assert len(words) % per_block == 0
    in_addr = 0x00501000
    asm = Assembler()
    for i, w in enumerate(words):
        asm.word(in_addr + 4 * i, w)
    if state is not None:
        for i, w in enumerate(state):
            asm.word(P2_GUEST_STATE_ADDR + 4 * i, w)
    if expect is not None:
        for i, w in enumerate(expect):
            asm.word(P2_GUEST_OUT_ADDR + 4 * i, w)
    a4, t4 = 14, 29
    asm.addi(a4, 0, 3)
    asm.li(REG_A0, P2_GUEST_STATE_ADDR if state is not None else 0)
    asm.li(REG_A1, in_addr)
    asm.li(REG_A2, P2_GUEST_OUT_ADDR)
    asm.li(REG_A3, (len(words) // per_block) | (PFLAG_IS_ELEM if is_elem else 0) | (PFLAG_CHECK_OUT if expect is not None else 0))
    asm.li(REG_A7, HOST_ECALL_POSEIDON2)
    asm.ecall()
    asm.li(t4, P2_GUEST_OUT_ADDR)
    asm.load(2, a4, t4, 0)
    asm.host_terminate(0, 0)
    entry, image = asm.program()
    return MemoryImage.new_kernel(entry, image)


def user_sha2_via_kernel_guest(message=b"abc"):
    """the production shape of an accelerator call (zkos/v1compat/src/kernel.s:60-100,190-240): a USER-mode program puts
    the arguments in its registers and executes `ecall`; the machine-mode kernel's dispatch handler reads them out of the
    user register file (plain memory at USER_REGS_ADDR), issues the machine ecall - sha2 here - and `mret`s back; a second
    user ecall with a7 = 0 makes the kernel terminate. The digest lands at SHA2_GUEST_OUT_ADDR."""
    t0, t1 = 5, 6
    data_addr, state_addr, k_addr = 0x00500000, 0x00500300, 0x00500500
    ml = len(message)
    padded = bytes(message) + b"\x80" + b"\x00" * ((55 - ml) % 64) + (8 * ml).to_bytes(8, "big")
    user = Assembler()
    for i in range(len(padded) // 4):
        user.word(data_addr + 4 * i, int.from_bytes(padded[4 * i:4 * i + 4], "little"))
    for i, w in enumerate(SHA256_IV):
        user.word(state_addr + 4 * i, _bswap(w))
    for i, w in enumerate(SHA256_K):
        user.word(k_addr + 4 * i, w)
    user.li(REG_A0, state_addr)
    user.li(REG_A1, SHA2_GUEST_OUT_ADDR)
    user.li(REG_A2, data_addr)
    user.li(REG_A3, len(padded) // 64)
    user.li(REG_A4, k_addr)
    user.li(REG_A7, 1)                  # "hash"
    user.ecall()
    user.li(t0, SHA2_GUEST_OUT_ADDR)
    user.load(2, t1, t0, 0)             # back in user mode: read the digest
    user.li(REG_A7, 0)                  # "exit"
    user.ecall()
    uentry, uimage = user.program()
    kern = Assembler(base=KERNEL_START_ADDR)
    kern.li(t1, uentry - 4)             # mret resumes at MEPC + 4
    kern.li(t0, MEPC_ADDR)
    kern.sw(t1, t0, 0)
    kern.mret()
    handler = KERNEL_START_ADDR + 0x100
    while KERNEL_START_ADDR + 4 * len(kern.text) < handler:
        kern.text.append(0x00000013)
    kern.li(t0, USER_REGS_ADDR)
    kern.load(2, t1, t0, 4 * REG_A7)
    kern.beq(t1, 0, 4 * 9)              # a7 == 0 -> the terminate sequence after the 8 instructions below
    for r in (REG_A0, REG_A1, REG_A2, REG_A3, REG_A4):
        kern.load(2, r, t0, 4 * r)
    kern.addi(REG_A7, 0, HOST_ECALL_SHA2)
    kern.ecall()
    kern.mret()
    kern.host_terminate(0, 0)
    kentry, kimage = kern.program()
    image = dict(uimage)
    image.update(kimage)
    image[ECALL_DISPATCH_ADDR] = handler
    return MemoryImage.new_kernel(kentry, image)
