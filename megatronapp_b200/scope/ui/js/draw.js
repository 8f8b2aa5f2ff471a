// Canvas primitives shared by the components: colour maps, matrix blit with hover read-out, 2-component PCA.
export function diverging(v) { const a = Math.min(1, Math.abs(v)); return v >= 0 ? [118 * a, 185 * a, 0] : [40 * a, 110 * a, 255 * a]; }
export function sequential(v) { const a = Math.max(0, Math.min(1, v)); return [255 * a, 200 * a * a, 40 * a]; }
const tip = document.createElement("div"); tip.className = "tip"; document.addEventListener("DOMContentLoaded", () => document.body.appendChild(tip));
export function drawMatrix(cv, rows, cols, get, opts = {}) {
  const ctx2 = cv.getContext("2d"); ctx2.clearRect(0, 0, cv.width, cv.height);
  if (!rows || !cols) return;
  const off = document.createElement("canvas"); off.width = cols; off.height = rows;
  const ctx = off.getContext("2d"), img = ctx.createImageData(cols, rows);
  let mx = opts.max || 1e-12; if (!opts.max) for (let r = 0; r < rows; r++) for (let c = 0; c < cols; c++) mx = Math.max(mx, Math.abs(get(r, c)));
  for (let r = 0; r < rows; r++) for (let c = 0; c < cols; c++) {
    const rgb = (opts.seq ? sequential : diverging)(get(r, c) / mx), i = 4 * (r * cols + c);
    img.data[i] = rgb[0]; img.data[i + 1] = rgb[1]; img.data[i + 2] = rgb[2]; img.data[i + 3] = 255;
  }
  ctx.putImageData(img, 0, 0); ctx2.imageSmoothingEnabled = false; ctx2.drawImage(off, 0, 0, cv.width, cv.height);
  if (opts.highlightRow !== undefined && opts.highlightRow >= 0 && opts.highlightRow < rows) {
    ctx2.strokeStyle = "#76b900"; ctx2.lineWidth = 1; const h = cv.height / rows; ctx2.strokeRect(0.5, opts.highlightRow * h + 0.5, cv.width - 1, Math.max(1, h - 1)); }
  cv.onmousemove = ev => { const b = cv.getBoundingClientRect(), r = Math.floor((ev.clientY - b.top) / b.height * rows), c = Math.floor((ev.clientX - b.left) / b.width * cols);
    if (r < 0 || c < 0 || r >= rows || c >= cols) return; tip.style.display = "block"; tip.style.left = ev.clientX + 12 + "px"; tip.style.top = ev.clientY + 12 + "px";
    tip.textContent = (opts.label ? opts.label(r, c) + "  " : `[${r}, ${c}]  `) + (+get(r, c)).toPrecision(4); };
  cv.onmouseleave = () => { tip.style.display = "none"; };
  if (opts.onPick) cv.onclick = ev => { const b = cv.getBoundingClientRect(); opts.onPick(Math.floor((ev.clientY - b.top) / b.height * rows), Math.floor((ev.clientX - b.left) / b.width * cols)); };
}
export function legend(cv, seq) { const c = cv.getContext("2d"); cv.width = 120; cv.height = 8;
  for (let x = 0; x < 120; x++) { const v = seq ? x / 119 : x / 59.5 - 1, rgb = (seq ? sequential : diverging)(v); c.fillStyle = `rgb(${rgb.map(Math.round)})`; c.fillRect(x, 0, 1, 8); } }
// two principal components by power iteration with deflation (the reference uses ml-pca; no dependency here)
export function pca2(X) {
  const n = X.length, d = X[0].length, mean = new Array(d).fill(0);
  X.forEach(r => r.forEach((v, j) => mean[j] += v / n));
  const Z = X.map(r => r.map((v, j) => v - mean[j])), comps = [], vars = [];
  for (let k = 0; k < 2; k++) {
    let w = Array.from({length:d}, (_, j) => Math.sin(j + 1 + k)), lam = 0;
    for (let it = 0; it < 80; it++) {
      const s = Z.map(r => r.reduce((a, v, j) => a + v * w[j], 0)), nw = new Array(d).fill(0);
      Z.forEach((r, i) => r.forEach((v, j) => nw[j] += v * s[i]));
      comps.forEach(c => { const dot = nw.reduce((a, v, j) => a + v * c[j], 0); nw.forEach((_, j) => nw[j] -= dot * c[j]); });
      lam = Math.hypot(...nw) || 1; w = nw.map(v => v / lam);
    }
    comps.push(w); vars.push(lam / Math.max(1, n - 1));
  }
  const total = Z.reduce((a, r) => a + r.reduce((b, v) => b + v * v, 0), 0) / Math.max(1, n - 1);
  return {points: Z.map(r => comps.map(c => r.reduce((a, v, j) => a + v * c[j], 0))), explained: vars.map(v => v / (total || 1))};
}
