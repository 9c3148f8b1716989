// swift-tools-version:5.9
// UNCOMPILED (see README.md): SwiftPM wiring of the C ABI module and the facade.
import PackageDescription

let package = Package(
  name: "SwiftMP3",
  products: [.library(name: "SwiftMP3", targets: ["SwiftMP3"])],
  targets: [
    .systemLibrary(name: "CMP3B200", path: "Sources/CMP3B200"),
    .target(name: "SwiftMP3", dependencies: ["CMP3B200"]),
  ]
)
