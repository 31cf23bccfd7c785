// grayshift_b200/host/image_io.cpp -- host-side image decoding (the reference uses the `image` and `radiant`
// crates: texture.rs:79, main.rs:139).  Decoded pixels travel to the device through gs_image.
#include <cmath>
#include <cstring>
#include <fstream>
#include <stdexcept>

#include "grayshift.hpp"

namespace grayshift {
namespace {

std::vector<uint8_t> read_all(const std::string& path) {
	std::ifstream f(path, std::ios::binary);
	if (!f) throw std::runtime_error("cannot open " + path);
	return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
bool ends_with(const std::string& s, const std::string& suf) {
	if (s.size() < suf.size()) return false;
	for (size_t i = 0; i < suf.size(); ++i) if (std::tolower((unsigned char)s[s.size() - suf.size() + i]) != suf[i]) return false;
	return true;
}
bool exists(const std::string& p) { std::ifstream f(p, std::ios::binary); return (bool)f; }

std::shared_ptr<ImageData> load_ppm(const std::string& path) {
	auto data = read_all(path);
	size_t pos = 0;
	auto token = [&]() {
		for (;;) {
			while (pos < data.size() && std::isspace(data[pos])) ++pos;
			if (pos < data.size() && data[pos] == '#') { while (pos < data.size() && data[pos] != '\n') ++pos; continue; }
			break;
		}
		std::string s;
		while (pos < data.size() && !std::isspace(data[pos])) s.push_back((char)data[pos++]);
		return s;
	};
	if (token() != "P6") throw std::runtime_error(path + ": not a binary PPM (P6)");
	auto img = std::make_shared<ImageData>();
	img->width = std::stoi(token()); img->height = std::stoi(token());
	if (std::stoi(token()) != 255) throw std::runtime_error(path + ": maxval must be 255");
	++pos;
	size_t n = (size_t)img->width * img->height * 3;
	if (img->width <= 0 || img->height <= 0 || data.size() - pos < n) throw std::runtime_error(path + ": truncated");
	img->format = GS_IMG_U8_RGB;
	img->u8.assign(data.begin() + pos, data.begin() + pos + n);
	img->source = "ppm:" + path;
	return img;
}

}  // namespace

// image::open(filepath) (texture.rs:79).  JPEG entropy decoding is not reimplemented here: a `.jpg`
// resolves to its decoded `.ppm` sibling (written once by grayshift_b200/assets.py with PIL; in the Rust
// shell the `image` crate decodes directly).
std::shared_ptr<ImageData> load_image_u8(const std::string& path) {
	if (ends_with(path, ".ppm")) return load_ppm(path);
	if (ends_with(path, ".jpg") || ends_with(path, ".jpeg")) {
		std::string stem = path.substr(0, path.rfind('.'));
		if (exists(stem + ".ppm")) return load_ppm(stem + ".ppm");
		throw std::runtime_error(path + ": no decoded sibling " + stem + ".ppm (run `python -m grayshift_b200.assets`)");
	}
	throw std::runtime_error(path + ": unsupported image format");
}

std::shared_ptr<ImageData> load_raw(const std::string& path, int w, int h, int format) {
	auto data = read_all(path);
	auto img = std::make_shared<ImageData>();
	img->width = w; img->height = h; img->format = format; img->source = "raw:" + path;
	size_t n = (size_t)w * h * 3;
	if (format == GS_IMG_U8_RGB) {
		if (data.size() != n) throw std::runtime_error(path + ": size mismatch");
		img->u8 = std::move(data);
	} else {
		if (data.size() != n * 4) throw std::runtime_error(path + ": size mismatch");
		img->f32.resize(n);
		std::memcpy(img->f32.data(), data.data(), data.size());
	}
	return img;
}

// radiant::load (main.rs:139, 806): Radiance RGBE, "-Y h +X w", new-style per-channel RLE or flat
// scanlines; RGB = mantissa * 2^(exponent - 136), exponent 0 -> black.
std::shared_ptr<ImageData> load_hdr(const std::string& path) {
	auto data = read_all(path);
	size_t pos = 0;
	auto line = [&]() { std::string s; while (pos < data.size() && data[pos] != '\n') s.push_back((char)data[pos++]); ++pos; return s; };
	if (line().rfind("#?", 0) != 0) throw std::runtime_error(path + ": not a Radiance picture");
	for (;;) {
		if (pos >= data.size()) throw std::runtime_error(path + ": header never ends");
		if (line().empty()) break;
	}
	int W = 0, H = 0;
	std::string res = line();
	if (std::sscanf(res.c_str(), "-Y %d +X %d", &H, &W) != 2 || W <= 0 || H <= 0) throw std::runtime_error(path + ": unsupported resolution line '" + res + "'");
	auto img = std::make_shared<ImageData>();
	img->width = W; img->height = H; img->format = GS_IMG_F32_RGB; img->source = "hdr:" + path;
	img->f32.resize((size_t)W * H * 3);
	std::vector<uint8_t> scan((size_t)W * 4);
	for (int y = 0; y < H; ++y) {
		if (pos + 4 > data.size()) throw std::runtime_error(path + ": truncated");
		const bool rle = W >= 8 && W < 32768 && data[pos] == 2 && data[pos + 1] == 2 && (((int)data[pos + 2] << 8) | data[pos + 3]) == W;
		if (rle) {
			pos += 4;
			for (int ch = 0; ch < 4; ++ch) {
				int x = 0;
				while (x < W) {
					if (pos >= data.size()) throw std::runtime_error(path + ": truncated");
					int count = data[pos++];
					if (count > 128) {
						count -= 128;
						if (x + count > W || pos >= data.size()) throw std::runtime_error(path + ": bad run");
						const uint8_t v = data[pos++];
						while (count--) scan[(size_t)(x++) * 4 + ch] = v;
					} else {
						if (count == 0 || x + count > W || pos + count > data.size()) throw std::runtime_error(path + ": bad literal run");
						while (count--) scan[(size_t)(x++) * 4 + ch] = data[pos++];
					}
				}
			}
		} else {
			if (pos + (size_t)W * 4 > data.size()) throw std::runtime_error(path + ": truncated");
			std::memcpy(scan.data(), &data[pos], (size_t)W * 4);
			pos += (size_t)W * 4;
		}
		float* o = &img->f32[(size_t)y * W * 3];
		for (int x = 0; x < W; ++x, o += 3) {
			const uint8_t* p = &scan[(size_t)x * 4];
			if (p[3] == 0) { o[0] = o[1] = o[2] = 0.f; continue; }
			const float f = std::ldexp(1.0f, (int)p[3] - 136);
			o[0] = (float)p[0] * f; o[1] = (float)p[1] * f; o[2] = (float)p[2] * f;
		}
	}
	return img;
}

}  // namespace grayshift
