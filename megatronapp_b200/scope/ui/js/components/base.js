// Minimal component base: a DOM root, a render() that runs when one of the subscribed store topics fires.
import {on} from "../store.js";
export class Component {
  constructor(tag = "div", cls = "") { this.el = document.createElement(tag); if (cls) this.el.className = cls; this._pending = false; }
  watch(...topics) { topics.forEach(t => on(t, () => this.schedule())); return this; }
  schedule() { if (this._pending) return; this._pending = true; requestAnimationFrame(() => { this._pending = false; this.render(); }); }
  render() {}
  mount(parent) { parent.appendChild(this.el); this.render(); return this; }
}
export function h(tag, attrs = {}, ...kids) {
  const e = document.createElement(tag);
  for (const [k, v] of Object.entries(attrs)) { if (k === "class") e.className = v; else if (k.startsWith("on")) e[k] = v; else if (v !== false && v != null) e.setAttribute(k, v); }
  kids.flat().forEach(c => e.appendChild(typeof c === "string" ? document.createTextNode(c) : c));
  return e;
}
export function panel(title, ...kids) { const p = h("div", {class:"panel"}, h("h2", {}, title), ...kids); return p; }
