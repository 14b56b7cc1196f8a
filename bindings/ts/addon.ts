// Loads the N-API addon (bindings/napi) and maps its status codes onto the reference's error classes
// (lib/StarkError.ts:3-13; TypeError for bad arguments as in lib/Stark.ts:86-87,172,322-340).
import { StarkError } from '../../lib/StarkError';

// eslint-disable-next-line @typescript-eslint/no-var-requires
export const native: any = require('../napi/build/Release/genstark_b200.node');

export const GS_E_ARG = -2, GS_E_UNSUPPORTED = -3, GS_E_STARK = -4;
export const HASH_IDS: { [alg: string]: number } = { sha256: 0, blake2s256: 1 };

/** runs fn; a negative status from the library arrives as an Error with .code -- rethrow it as the class genSTARK throws */
export function guarded<T>(fn: () => T): T {
    try { return fn(); }
    catch (e) {
        if (e && e.code === GS_E_STARK) throw new StarkError(e.message);
        if (e && e.code === GS_E_ARG) throw new TypeError(e.message);
        throw e;
    }
}

/** field element <-> the 16 little-endian bytes Vector.toBuffer() yields (lib/utils/serialization.ts:131-147) */
export function toBytes16(v: bigint): Buffer {
    const b = Buffer.alloc(16);
    b.writeBigUInt64LE(v & 0xFFFFFFFFFFFFFFFFn, 0); b.writeBigUInt64LE(v >> 64n, 8);
    return b;
}
export function fromBytes16(b: Buffer, offset = 0): bigint {
    return b.readBigUInt64LE(offset) | (b.readBigUInt64LE(offset + 8) << 64n);
}
export function packElements(values: bigint[]): Buffer {
    return Buffer.concat(values.map(toBytes16));
}
