// UNCOMPILED facade over libmp3b200.so (no Swift toolchain in the build image; see ../../README.md).
// Same public surface as the reference's Sources/SwiftMP3/MP3Encoder.swift ("SRC"): ID3Tag (SRC:8-54), MP3EncoderOptions
// (SRC:57-116, same init labels and defaults), MP3Encoder (SRC:132-230: newSession(), async encode(_:), encode(_:to:)),
// EncoderSession (SRC:237-350: encode(samples:), flush(), generateXingHeader(), generateID3Tag(), encodedFrameCount,
// encodedByteCount).  EncoderSession stays a copyable struct with mutating methods like the reference's: the GPU-side state
// lives behind a reference-counted box and is cloned lazily (mp3b_session_clone) when a copy is mutated, so
// `var fork = session` keeps the reference's value semantics.
import CMP3B200
import Foundation

public struct ID3Tag: Sendable, Equatable {
  public var title: String?, artist: String?, album: String?
  public var track: UInt16?, trackTotal: UInt16?, year: UInt16?
  public var genre: String?, comment: String?
  public var albumArt: Data?
  public var albumArtMIMEType: String
  public init(title: String? = nil, artist: String? = nil, album: String? = nil, track: UInt16? = nil, trackTotal: UInt16? = nil,
              year: UInt16? = nil, genre: String? = nil, comment: String? = nil, albumArt: Data? = nil,
              albumArtMIMEType: String = "image/jpeg") {
    self.title = title; self.artist = artist; self.album = album; self.track = track; self.trackTotal = trackTotal
    self.year = year; self.genre = genre; self.comment = comment; self.albumArt = albumArt; self.albumArtMIMEType = albumArtMIMEType
  }
}

public struct MP3EncoderOptions: Sendable, Equatable {
  public enum Mode: String, Sendable, Equatable { case mono, stereo, jointStereo }
  public var sampleRate: Int, bitrateKbps: Int
  public var vbr: Bool
  public var mode: Mode
  public var quality: Int
  public var crcProtected: Bool, original: Bool, copyright: Bool
  public var id3Tag: ID3Tag?
  public init(sampleRate: Int = 44_100, bitrateKbps: Int = 128, vbr: Bool = false, mode: Mode = .stereo, quality: Int = 5,
              crcProtected: Bool = false, original: Bool = true, copyright: Bool = false, id3Tag: ID3Tag? = nil) {
    self.sampleRate = sampleRate; self.bitrateKbps = bitrateKbps; self.vbr = vbr; self.mode = mode
    self.quality = max(0, min(quality, 9))                                   // SRC:110
    self.crcProtected = crcProtected; self.original = original; self.copyright = copyright; self.id3Tag = id3Tag
  }

  fileprivate var cOptions: mp3b_options {
    var o = mp3b_options()
    mp3b_options_default(&o)
    o.sample_rate = Int32(sampleRate); o.bitrate_kbps = Int32(bitrateKbps); o.vbr = vbr ? 1 : 0
    o.mode = mode == .mono ? 0 : (mode == .stereo ? 1 : 2)
    o.quality = Int32(quality); o.crc_protected = crcProtected ? 1 : 0
    o.original = original ? 1 : 0; o.copyright = copyright ? 1 : 0
    return o
  }
}

public struct MP3Encoder: Sendable {
  public let options: MP3EncoderOptions
  /// CUDA ordinal new sessions are created on (an addition; the reference has no device).
  public var device: Int32 = 0
  public init(options: MP3EncoderOptions = MP3EncoderOptions()) { self.options = options }

  public func newSession() -> EncoderSession { EncoderSession(options: options, device: device) }       // SRC:143-145

  /// SRC:151-179: frames as they are produced; no Xing header.
  public func encode<S: AsyncSequence & Sendable>(_ input: S) -> AsyncThrowingStream<Data, Error> where S.Element == [Float] {
    let encoder = self
    return AsyncThrowingStream { continuation in
      let task = Task {
        var session = encoder.newSession()
        do {
          for try await samples in input {
            try Task.checkCancellation()
            let data = session.encode(samples: samples)
            if !data.isEmpty { continuation.yield(data) }
          }
          let tail = session.flush()
          if !tail.isEmpty { continuation.yield(tail) }
          continuation.finish()
        } catch { continuation.finish(throwing: error) }
      }
      continuation.onTermination = { _ in task.cancel() }
    }
  }

  /// SRC:189-230: [ID3 tag][Xing placeholder][frames...], then the real Xing frame over the placeholder.
  public func encode<S: AsyncSequence & Sendable>(_ input: S, to url: URL) async throws where S.Element == [Float] {
    var session = newSession()
    let id3 = session.generateID3Tag()
    var o = options.cOptions
    let placeholder = Int(mp3b_xing_frame_size(&o))                          // SRC:198-200: from the snapped bitrate
    var head = id3
    head.append(Data(count: placeholder))
    try head.write(to: url)
    let file = try FileHandle(forWritingTo: url)
    defer { try? file.close() }
    try file.seek(toOffset: UInt64(id3.count + placeholder))
    for try await samples in input {
      try Task.checkCancellation()
      let data = session.encode(samples: samples)
      if !data.isEmpty { try file.write(contentsOf: data) }
    }
    let tail = session.flush()
    if !tail.isEmpty { try file.write(contentsOf: tail) }
    try file.seek(toOffset: UInt64(id3.count))
    try file.write(contentsOf: session.generateXingHeader())
  }
}

/// Owner of one mp3b_session handle.  Not Sendable, like the reference's session (README.md:207: one context at a time).
final class SessionBox {
  let handle: OpaquePointer
  init(options: MP3EncoderOptions, device: Int32) {
    var o = options.cOptions
    var h: OpaquePointer?
    precondition(mp3b_session_create(&o, device, &h) == 0, String(cString: mp3b_last_error()))
    handle = h!
  }
  init(cloning other: SessionBox) {
    var h: OpaquePointer?
    precondition(mp3b_session_clone(other.handle, &h) == 0, String(cString: mp3b_last_error()))
    handle = h!
  }
  deinit { mp3b_session_destroy(handle) }
}

public struct EncoderSession {
  private var box: SessionBox
  private let options: MP3EncoderOptions

  init(options: MP3EncoderOptions, device: Int32 = 0) { self.options = options; box = SessionBox(options: options, device: device) }

  /// Value semantics of the reference's struct (SRC:237-258): a copy that is mutated gets its own snapshot of the encoder.
  private mutating func makeUnique() { if !isKnownUniquelyReferenced(&box) { box = SessionBox(cloning: box) } }

  public var encodedFrameCount: UInt32 { mp3b_session_frame_count(box.handle) }   // SRC:261
  public var encodedByteCount: UInt32 { mp3b_session_byte_count(box.handle) }     // SRC:264

  /// Opt-in ISO mode (include/mp3b200.h); call before the first encode(samples:).
  /// Extension (not in the reference; default 0 = the reference's bytes): 1 = ISO quantizer / table selection / count1 / real
  /// main_data_begin, 2 = + psychoacoustic model and scalefactor outer loop, 3 = + window switching.  Fresh sessions only.
  public mutating func setISOMode(_ level: Int32) { makeUnique(); precondition(mp3b_session_set_iso_mode(box.handle, level) == 0, String(cString: mp3b_last_error())) }
  public mutating func setISOMode(_ on: Bool) { setISOMode(on ? 1 : 0) }

  private func collect(_ call: (UnsafeMutablePointer<UInt8>?, Int, UnsafeMutablePointer<Int>) -> Int32, bound: Int) -> Data {
    var out = Data(count: bound)
    var written = 0
    var rc = out.withUnsafeMutableBytes { buf in call(buf.bindMemory(to: UInt8.self).baseAddress, buf.count, &written) }
    if rc == MP3B_ERR_BUFFER_TOO_SMALL.rawValue {                            // nothing is lost: the frames wait in the handle
      out = Data(count: written)
      rc = out.withUnsafeMutableBytes { buf in mp3b_session_take_output(box.handle, buf.bindMemory(to: UInt8.self).baseAddress, buf.count, &written) }
    }
    precondition(rc == 0, String(cString: mp3b_last_error()))                // the reference's sync API cannot fail either
    out.removeSubrange(written...)
    return out
  }

  public mutating func encode(samples: [Float]) -> Data {                    // SRC:297-310
    makeUnique()
    let h = box.handle
    return samples.withUnsafeBufferPointer { pcm in
      collect({ mp3b_session_encode(h, pcm.baseAddress, samples.count, $0, $1, $2) }, bound: mp3b_session_output_bound(h, samples.count))
    }
  }

  public mutating func flush() -> Data {                                     // SRC:318-350
    makeUnique()
    let h = box.handle
    return collect({ mp3b_session_flush(h, $0, $1, $2) }, bound: mp3b_session_output_bound(h, 0))
  }

  public func generateXingHeader() -> Data {                                 // SRC:367-449
    let h = box.handle
    return collect({ mp3b_session_xing_header(h, $0, $1, $2) }, bound: 2048)
  }

  public func generateID3Tag() -> Data {                                     // SRC:355-358
    guard let tag = options.id3Tag else { return Data() }
    func withOptionalCString<R>(_ s: String?, _ body: (UnsafePointer<CChar>?) -> R) -> R {
      guard let s else { return body(nil) }
      return s.withCString { body($0) }
    }
    let art = tag.albumArt.map { [UInt8]($0) }
    return withOptionalCString(tag.title) { title in withOptionalCString(tag.artist) { artist in withOptionalCString(tag.album) { album in
      withOptionalCString(tag.genre) { genre in withOptionalCString(tag.comment) { comment in tag.albumArtMIMEType.withCString { mime in
        (art ?? []).withUnsafeBufferPointer { artBytes in
          var t = mp3b_id3(title: title, artist: artist, album: album, genre: genre, comment: comment,
                           track: tag.track.map(Int32.init) ?? -1, track_total: tag.trackTotal.map(Int32.init) ?? -1,
                           year: tag.year.map(Int32.init) ?? -1, album_art: art == nil ? nil : artBytes.baseAddress,
                           album_art_len: art?.count ?? 0, album_art_mime: mime)
          var need = 0
          _ = mp3b_id3_build(&t, nil, 0, &need)                              // size first (an empty tag needs 0 bytes, SRC:1066)
          var out = Data(count: need)
          if need > 0 { _ = out.withUnsafeMutableBytes { mp3b_id3_build(&t, $0.bindMemory(to: UInt8.self).baseAddress, need, &need) } }
          return out
        } } } } } } }
  }
}
