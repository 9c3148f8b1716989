// UNCOMPILED facade over libmp3b200.so (see ../../README.md).  Replaces the stored properties and bodies of the
// reference's EncoderSession (Sources/SwiftMP3/MP3Encoder.swift, lines 237-350); MP3EncoderOptions and the async
// conveniences of MP3Encoder (lines 57-230) keep their reference source unchanged, only `newSession()` is shown.
import CMP3B200
import Foundation

public struct MP3EncoderOptions: Sendable {
  public enum Mode: Sendable { case mono, stereo, jointStereo }
  public var sampleRate = 44_100, bitrateKbps = 128
  public var vbr = false
  public var mode: Mode = .stereo
  public var quality = 5 { didSet { quality = min(max(quality, 0), 9) } }
  public var crcProtected = false, original = true, copyright = false
  public init() {}
}

public struct MP3Encoder: Sendable {
  public let options: MP3EncoderOptions
  public init(options: MP3EncoderOptions = MP3EncoderOptions()) { self.options = options }
  public func newSession(device: Int32 = 0) -> EncoderSession { EncoderSession(options: options, device: device) }
}

public final class EncoderSession {      // a class (the handle owns GPU memory); `copy()` is the reference's struct copy
  private let handle: OpaquePointer
  public let options: MP3EncoderOptions

  init(options: MP3EncoderOptions, device: Int32) {
    var o = mp3b_options()
    mp3b_options_default(&o)
    o.sample_rate = Int32(options.sampleRate); o.bitrate_kbps = Int32(options.bitrateKbps)
    o.vbr = options.vbr ? 1 : 0
    o.mode = options.mode == .mono ? 0 : (options.mode == .stereo ? 1 : 2)
    o.quality = Int32(options.quality); o.crc_protected = options.crcProtected ? 1 : 0
    o.original = options.original ? 1 : 0; o.copyright = options.copyright ? 1 : 0
    var h: OpaquePointer?
    precondition(mp3b_session_create(&o, device, &h) == 0, String(cString: mp3b_last_error()))
    handle = h!; self.options = options
  }
  private init(handle: OpaquePointer, options: MP3EncoderOptions) { self.handle = handle; self.options = options }
  deinit { mp3b_session_destroy(handle) }

  /// `var copy = session` of the reference: an independent snapshot of the encoder.
  public func copy() -> EncoderSession {
    var h: OpaquePointer?
    precondition(mp3b_session_clone(handle, &h) == 0, String(cString: mp3b_last_error()))
    return EncoderSession(handle: h!, options: options)
  }

  public var encodedFrameCount: UInt32 { mp3b_session_frame_count(handle) }
  public var encodedByteCount: UInt32 { mp3b_session_byte_count(handle) }

  public func encode(samples: [Float]) -> Data {
    var out = Data(count: mp3b_session_output_bound(handle, samples.count))
    var written = 0
    let rc = samples.withUnsafeBufferPointer { pcm in
      out.withUnsafeMutableBytes { buf in
        mp3b_session_encode(handle, pcm.baseAddress, samples.count, buf.bindMemory(to: UInt8.self).baseAddress, buf.count, &written)
      }
    }
    precondition(rc == 0, String(cString: mp3b_last_error()))
    out.removeSubrange(written...)
    return out
  }

  public func flush() -> Data {
    var out = Data(count: mp3b_session_output_bound(handle, 0))
    var written = 0
    let rc = out.withUnsafeMutableBytes { buf in
      mp3b_session_flush(handle, buf.bindMemory(to: UInt8.self).baseAddress, buf.count, &written)
    }
    precondition(rc == 0, String(cString: mp3b_last_error()))
    out.removeSubrange(written...)
    return out
  }

  public func generateXingHeader() -> Data {
    var out = Data(count: 2048)
    var written = 0
    _ = out.withUnsafeMutableBytes { mp3b_session_xing_header(handle, $0.bindMemory(to: UInt8.self).baseAddress, 2048, &written) }
    out.removeSubrange(written...)
    return out
  }
}
