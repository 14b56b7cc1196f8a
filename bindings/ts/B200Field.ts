// A third implementation of galois' FiniteField next to "JS bigint" and "WASM": vectors and matrices live in HBM behind
// gs_mat handles, every O(n) method is one call into libgenstark_b200.so.  Method names, argument meaning and inv(0) = 0
// are the reference's (call sites: SURVEY.md section 8b; e.g. lib/Stark.ts:106,109, lib/components/CompositionPolynomial.ts:94-145,
// LinearCombination.ts:36-64, LowDegreeProver.ts:42-53,137-140,176-221, BoundaryConstraints.ts:71-95, ZeroPolynomial.ts:36-45).
import { native, guarded, toBytes16, fromBytes16, packElements } from './addon';

const P128 = 2n ** 128n - 9n * 2n ** 32n + 1n;
const registry = new FinalizationRegistry<any>(handle => native.matFree(handle));

export class B200Matrix {
    constructor(readonly ctx: any, readonly handle: any) { registry.register(this, handle); }
    private shape(): [number, number] { return native.matShape(this.handle); }
    get rowCount(): number { return this.shape()[0]; }
    get colCount(): number { return this.shape()[1]; }
    get length(): number { const [r, c] = this.shape(); return r === 1 ? c : r; }
    get elementSize(): number { return 16; }
    toBuffer(): Buffer {
        const [r, c] = this.shape();
        const out = Buffer.alloc(16 * r * c);
        guarded(() => native.matToBytes(this.ctx, this.handle, out));
        return out;
    }
    toValues(): bigint[] | bigint[][] {
        const [r, c] = this.shape(), raw = this.toBuffer();
        const row = (i: number) => Array.from({ length: c }, (_, j) => fromBytes16(raw, 16 * (i * c + j)));
        return r === 1 ? row(0) : Array.from({ length: r }, (_, i) => row(i));
    }
    getValue(row: number, column?: number): bigint {
        const [r] = this.shape();
        const [i, j] = column === undefined ? (r === 1 ? [0, row] : [row, 0]) : [row, column];
        return fromBytes16(guarded(() => native.matGet(this.ctx, this.handle, i, j)));
    }
    copyValue(index: number, destination: Buffer, offset: number): number {
        guarded(() => native.matGet(this.ctx, this.handle, 0, index)).copy(destination, offset);
        return 16;
    }
    rowsToBuffers(indexes?: number[]): Buffer[] {
        const [r, c] = this.shape(), raw = this.toBuffer();
        return (indexes || Array.from({ length: r }, (_, i) => i)).map(i => raw.subarray(16 * i * c, 16 * (i + 1) * c));
    }
}

export class B200Field {
    readonly characteristic = P128;
    readonly extensionDegree = 1;
    readonly elementSize = 16;
    readonly zero = 0n;
    readonly one = 1n;
    readonly isOptimized = true;                       // lib/Stark.ts:41-43 reads this to decide about the warning

    constructor(readonly ctx: any) {
        if (native.fieldSupported(toBytes16(P128)) !== 0) throw new TypeError('field is not supported by the B200 backend');
    }
    static isSupported(modulus: bigint): boolean { return modulus < (1n << 128n) && native.fieldSupported(toBytes16(modulus)) === 0; }

    private wrap(handle: any): B200Matrix { return new B200Matrix(this.ctx, handle); }
    private scalar(op: number, a: bigint, b: bigint): bigint { return fromBytes16(guarded(() => native.fieldScalarOp(op, toBytes16(a), toBytes16(b)))); }
    private mod(v: bigint): bigint { const m = v % P128; return m < 0n ? m + P128 : m; }

    // scalars -----------------------------------------------------------------------------------
    add(a: bigint, b: bigint): bigint { return this.scalar(0, this.mod(a), this.mod(b)); }
    sub(a: bigint, b: bigint): bigint { return this.scalar(1, this.mod(a), this.mod(b)); }
    mul(a: bigint, b: bigint): bigint { return this.scalar(2, this.mod(a), this.mod(b)); }
    div(a: bigint, b: bigint): bigint { return this.scalar(3, this.mod(a), this.mod(b)); }
    neg(a: bigint): bigint { return this.scalar(1, 0n, this.mod(a)); }
    inv(a: bigint): bigint { return this.scalar(3, 1n, this.mod(a)); }
    exp(base: bigint, exponent: bigint): bigint {
        if (exponent < 0n) { base = this.inv(base); exponent = -exponent; }         // galois: negative exponents invert
        let r = 1n, b = this.mod(base);
        for (let e = exponent; e > 0n; e >>= 1n) { if (e & 1n) r = this.mul(r, b); b = this.mul(b, b); }
        return r;
    }
    getRootOfUnity(order: number): bigint { return fromBytes16(guarded(() => native.fieldRootOfUnity(Math.log2(order)))); }
    prng(seed: bigint | Buffer, length?: number): bigint | B200Matrix {
        const s = Buffer.isBuffer(seed) ? seed : Buffer.from(seed.toString(16), 'hex');   // the odd-length hex quirk is the library's job
        const out = Buffer.alloc(16 * (length || 1));
        guarded(() => native.fieldPrng(s, length || 1, out));
        return length === undefined ? fromBytes16(out) : this.wrap(guarded(() => native.matFromBytes(this.ctx, out, 1, length)));
    }

    // constructors --------------------------------------------------------------------------------
    newVectorFrom(values: bigint[]): B200Matrix { return this.wrap(guarded(() => native.matFromBytes(this.ctx, packElements(values.map(v => this.mod(v))), 1, values.length))); }
    newMatrixFrom(rows: bigint[][]): B200Matrix {
        return this.wrap(guarded(() => native.matFromBytes(this.ctx, packElements(rows.flat().map(v => this.mod(v))), rows.length, rows[0].length)));
    }
    newMatrixFromVectors(vectors: B200Matrix[]): B200Matrix { return this.wrap(guarded(() => native.matStack(this.ctx, vectors.map(v => v.handle)))); }
    matrixRowsToVectors(m: B200Matrix): B200Matrix[] {
        return Array.from({ length: m.rowCount }, (_, i) => this.wrap(guarded(() => native.matRows(this.ctx, m.handle, i, 1))));
    }

    // vectors -----------------------------------------------------------------------------------
    private binary(op: number, a: B200Matrix, b: B200Matrix | bigint): B200Matrix {
        const h = typeof b === 'bigint' ? native.vecBinary(this.ctx, op, a.handle, null, toBytes16(this.mod(b)))
                                        : native.vecBinary(this.ctx, op, a.handle, b.handle, null);
        return this.wrap(h);
    }
    addVectorElements(a: B200Matrix, b: B200Matrix | bigint): B200Matrix { return guarded(() => this.binary(0, a, b)); }
    subVectorElements(a: B200Matrix, b: B200Matrix | bigint): B200Matrix { return guarded(() => this.binary(1, a, b)); }
    mulVectorElements(a: B200Matrix, b: B200Matrix | bigint): B200Matrix { return guarded(() => this.binary(2, a, b)); }
    divVectorElements(a: B200Matrix, b: B200Matrix | bigint): B200Matrix {
        if (typeof b === 'bigint') return this.mulVectorElements(a, this.inv(b));
        return this.wrap(guarded(() => native.vecDiv(this.ctx, a.handle, b.handle)));
    }
    subMatrixElementsFromVectors(vectors: B200Matrix[], m: B200Matrix): B200Matrix { return guarded(() => this.binary(1, this.newMatrixFromVectors(vectors), m)); }
    divMatrixElements(a: B200Matrix, b: B200Matrix): B200Matrix { return this.wrap(guarded(() => native.vecDiv(this.ctx, a.handle, b.handle))); }
    expVectorElements(a: B200Matrix, exponent: bigint): B200Matrix {
        if (exponent < 0n) return this.expVectorElements(this.divVectorElements(this.newVectorFrom(new Array(a.length).fill(1n)), a), -exponent);
        return this.wrap(guarded(() => native.vecExp(this.ctx, a.handle, toBytes16(exponent))));
    }
    mulMatrixByVector(m: B200Matrix, v: B200Matrix): B200Matrix { return this.wrap(guarded(() => native.matMulVector(this.ctx, m.handle, v.handle))); }
    combineVectors(a: B200Matrix, b: B200Matrix): bigint { return fromBytes16(guarded(() => native.vecCombine(this.ctx, a.handle, b.handle))); }
    combineManyVectors(vectors: B200Matrix[], coefficients: B200Matrix): B200Matrix {
        return this.wrap(guarded(() => native.vecCombineMany(this.ctx, vectors.map(v => v.handle), coefficients.toBuffer())));
    }
    getPowerSeries(base: bigint, length: number): B200Matrix { return this.wrap(guarded(() => native.powerSeries(this.ctx, toBytes16(this.mod(base)), length))); }
    pluckVector(v: B200Matrix, skip: number, times: number): B200Matrix { return this.wrap(guarded(() => native.pluckVector(this.ctx, v.handle, skip, times))); }
    transposeVector(v: B200Matrix, columns: number, step = 1): B200Matrix { return this.wrap(guarded(() => native.transposeVector(this.ctx, v.handle, columns, step))); }
    transposeMatrix(m: B200Matrix): B200Matrix { return this.wrap(guarded(() => native.matTranspose(this.ctx, m.handle))); }
    joinMatrixRows(m: B200Matrix): B200Matrix {
        const copy = this.wrap(guarded(() => native.matRows(this.ctx, m.handle, 0, m.rowCount)));
        guarded(() => native.matReshape(copy.handle, 1, m.rowCount * m.colCount));
        return copy;
    }

    // polynomials ---------------------------------------------------------------------------------
    interpolateRoots(_domain: B200Matrix, values: B200Matrix): B200Matrix { return this.wrap(guarded(() => native.interpolateRoots(this.ctx, values.handle))); }
    evalPolyAtRoots(poly: B200Matrix, domain: B200Matrix): B200Matrix { return this.evalPolysAtRoots(poly, domain); }
    evalPolysAtRoots(polys: B200Matrix, domain: B200Matrix): B200Matrix {
        return this.wrap(guarded(() => native.evalPolysAtRoots(this.ctx, polys.handle, Math.log2(domain.length))));
    }
    interpolate(xs: B200Matrix, ys: B200Matrix): B200Matrix {
        const n = xs.length, out = Buffer.alloc(16 * n);
        guarded(() => native.polyInterpolate(xs.toBuffer(), ys.toBuffer(), n, out));
        return this.wrap(native.matFromBytes(this.ctx, out, 1, n));
    }
    evalPolyAt(poly: B200Matrix, x: bigint): bigint { return fromBytes16(guarded(() => native.polyEvalAt(poly.toBuffer(), poly.length, toBytes16(this.mod(x))))); }
    mulPolys(a: B200Matrix, b: B200Matrix): B200Matrix {
        const out = Buffer.alloc(16 * (a.length + b.length - 1));
        guarded(() => native.polyMul(a.toBuffer(), a.length, b.toBuffer(), b.length, out));
        return this.wrap(native.matFromBytes(this.ctx, out, 1, a.length + b.length - 1));
    }
    interpolateQuarticBatch(xSets: B200Matrix, ySets: B200Matrix): B200Matrix { return this.wrap(guarded(() => native.quarticInterpolateBatch(this.ctx, xSets.handle, ySets.handle))); }
    evalQuarticBatch(polys: B200Matrix, x: bigint | B200Matrix): B200Matrix {
        const h = typeof x === 'bigint' ? native.quarticEvalBatch(this.ctx, polys.handle, null, toBytes16(this.mod(x)))
                                        : native.quarticEvalBatch(this.ctx, polys.handle, x.handle, null);
        return this.wrap(h);
    }
    /** interpolateQuarticBatch + evalQuarticBatch of one FRI layer in a single kernel (LowDegreeProver.ts:190-195) */
    friFold(v: B200Matrix, domainSize: number, depth: number, specialX: bigint): B200Matrix {
        return this.wrap(guarded(() => native.friFold(this.ctx, v.handle, Math.log2(domainSize), depth, toBytes16(specialX))));
    }
}
